"""CPU-only: the C++ LSAP (dskd_b200/csrc/lsap.cpp) returns exactly SciPy's indices -- the third-party
solver the reference calls at gfl_hungarian_assigner.py:147."""
import ctypes as C

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from dskd_b200 import _lib, lsap


def same(m):
    a, b = linear_sum_assignment(m)
    ri, ci = lsap(m)
    return np.array_equal(a, ri.numpy()) and np.array_equal(b, ci.numpy())


def test_random_tied_and_constant_matrices():
    rng = np.random.default_rng(0)
    for t in range(4000):
        r, c = rng.integers(1, 48), rng.integers(1, 48)
        kind = t % 5
        if kind == 0:
            m = rng.random((r, c))
        elif kind == 1:
            m = rng.integers(0, 4, (r, c)).astype(float)          # many ties
        elif kind == 2:
            m = np.full((r, c), 3.0)                              # constant (SciPy issue 11602 order)
        elif kind == 3:
            m = rng.standard_normal((r, c)).astype(np.float32).astype(float)
        else:
            m = np.round(rng.random((r, c)), 1)
        assert same(m), (t, r, c)


def test_detr_shapes_rows_gt_cols():
    rng = np.random.default_rng(1)
    for g in (1, 5, 37, 100, 300):
        assert same(rng.standard_normal((300, g)).astype(np.float32).astype(float))
    assert same(rng.random((37, 300)))


def test_empty_and_infeasible():
    ri, ci = lsap(np.zeros((0, 5)))
    assert ri.numel() == 0 and ci.numel() == 0
    with pytest.raises(ValueError):
        lsap(np.array([[np.nan, 1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError):
        linear_sum_assignment(np.array([[np.nan, 1.0], [1.0, 2.0]]))
    m = np.array([[np.inf, np.inf], [1.0, 2.0]])
    with pytest.raises(ValueError):
        lsap(m)
    with pytest.raises(ValueError):
        linear_sum_assignment(m)
    m = np.array([[np.inf, 1.0], [1.0, np.inf]])
    assert same(m)


def test_batch_f32_matches_per_problem_scipy():
    rng = np.random.default_rng(2)
    P, Q, ld = 12, 300, 64
    cols = rng.integers(0, ld + 1, P).astype(np.int32)
    cost = rng.standard_normal((P, Q, ld)).astype(np.float32)
    out = np.full((P, Q), -7, dtype=np.int64)
    for threads in (1, 4, 0):
        rc = _lib.load().dskd_lsap_batch_f32(C.c_void_p(cost.ctypes.data), P, Q, ld, C.c_void_p(cols.ctypes.data),
                                             C.c_void_p(out.ctypes.data), threads)
        assert rc == 0
        for p in range(P):
            ref = np.zeros(Q, dtype=np.int64)
            if cols[p]:
                r, c = linear_sum_assignment(torch.from_numpy(cost[p, :, :cols[p]].copy()))
                ref[r] = c + 1
            assert np.array_equal(out[p], ref), (threads, p)
