"""The device LSAP (`dskd_lsap_batch_device`) must return SciPy's indices bit for bit: random, tied, constant,
rectangular both ways, ragged batches, infeasible and NaN matrices."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from dskd_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def solve_device(mats, rows, ld):
    """mats: list of [rows, c_p] float32 arrays -> (assigned [P, rows] int64, status [P])."""
    P = len(mats)
    cost = torch.full((P, rows, ld), 7.5, dtype=torch.float32)          # padding is never read
    cols = []
    for p, m in enumerate(mats):
        cost[p, :, :m.shape[1]] = torch.from_numpy(m)
        cols.append(m.shape[1])
    start = torch.tensor([0] + np.cumsum(cols).tolist(), dtype=torch.int32, device=DEV)
    cost = cost.to(DEV)
    out = torch.empty(P, rows, dtype=torch.int64, device=DEV)
    status = torch.empty(P, dtype=torch.int32, device=DEV)
    lib = L.load()
    L.check(lib.dskd_lsap_batch_device(L.ptr(cost), P, P, rows, ld, L.ptr(start), max(cols + [0]), L.ptr(out), L.ptr(status),
                                       L.stream_of(cost)), 'dskd_lsap_batch_device')
    torch.cuda.synchronize()
    return out.cpu().numpy(), status.cpu().numpy()


def scipy_assigned(m):
    out = np.zeros(m.shape[0], dtype=np.int64)
    if m.shape[1]:
        r, c = linear_sum_assignment(m.astype(np.float64))
        out[r] = c + 1
    return out


@pytest.mark.parametrize('rows,max_cols', [(300, 50), (300, 110), (100, 100), (37, 64), (8, 300), (300, 1)])
def test_random_and_tied_batches_match_scipy(rows, max_cols):
    rng = np.random.RandomState(rows * 1000 + max_cols)
    mats = []
    for p in range(24):
        c = int(rng.randint(1, max_cols + 1)) if p else max_cols
        kind = p % 4
        if kind == 0:
            m = rng.randn(rows, c)
        elif kind == 1:
            m = rng.randint(0, 4, size=(rows, c)).astype(np.float64)          # heavy ties
        elif kind == 2:
            m = np.round(rng.rand(rows, c) * 8) / 8 - 2.0                      # ties + negative
        else:
            m = np.full((rows, c), 3.0)                                        # constant
        mats.append(m.astype(np.float32))
    mats.append(np.zeros((rows, 0), dtype=np.float32))                         # an image without GT
    got, status = solve_device(mats, rows, max_cols)
    assert (status == 0).all()
    for p, m in enumerate(mats):
        assert np.array_equal(got[p], scipy_assigned(m)), (p, m.shape)


def test_infeasible_and_nan_are_reported():
    rows = 6
    ok = np.arange(24, dtype=np.float32).reshape(6, 4)
    inf_col = ok.copy()
    inf_col[:, 2] = np.inf                           # a column nobody can take: infeasible (rows > cols, transposed)
    nan = ok.copy()
    nan[3, 1] = np.nan
    neg = ok.copy()
    neg[0, 0] = -np.inf
    some_inf = ok.copy()
    some_inf[0, :] = np.inf                          # one query unusable: still feasible
    got, status = solve_device([ok, inf_col, nan, neg, some_inf], rows, 4)
    assert status.tolist() == [0, L.EINFEASIBLE, L.EINFEASIBLE, L.EINFEASIBLE, 0]
    assert np.array_equal(got[0], scipy_assigned(ok)) and np.array_equal(got[4], scipy_assigned(some_inf))
    assert (got[1] == 0).all() and (got[2] == 0).all() and (got[3] == 0).all()
    with pytest.raises(ValueError):
        linear_sum_assignment(inf_col.astype(np.float64))


def test_detr_cost_matrices_device_equals_host_solver():
    import dskd_b200
    from dskd_b200 import synth
    ai = synth.make_assign_inputs(num_images=8, seed=5)
    args = (ai.cls_logits.to(DEV), ai.box_pred.to(DEV), [g.to(DEV) for g in ai.gt_bboxes], [l.to(DEV) for l in ai.gt_labels],
            ai.img_shapes)
    dev_res = dskd_b200.GFLHungarianAssigner(solver='device').assign_batch(*args, prev_labels=list(range(40)))
    host_res = dskd_b200.GFLHungarianAssigner(solver='host').assign_batch(*args, prev_labels=list(range(40)))
    for k in dev_res:
        assert torch.equal(dev_res[k].cpu(), host_res[k].cpu()), k
