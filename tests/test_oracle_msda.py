"""The deformable-attention oracle (oracle/msda.py: the op's grid_sample definition) against a literal per-tap loop of
the same definition -- pixel = loc * size - 0.5, four bilinear taps, taps outside the map contribute 0 -- which is the
arithmetic csrc/msda.cu implements.  CPU only, tiny sizes."""
import math

import torch

from oracle.msda import msda_torch


def _loop(value, shapes, loc, attn):
    N, S, M, D = value.shape
    Lq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    out = torch.zeros(N, Lq, M, D, dtype=value.dtype)
    starts = [0]
    for h, w in shapes:
        starts.append(starts[-1] + h * w)
    for n in range(N):
        for q in range(Lq):
            for m in range(M):
                for l, (H, W) in enumerate(shapes):
                    for p in range(P):
                        ix = float(loc[n, q, m, l, p, 0]) * W - 0.5
                        iy = float(loc[n, q, m, l, p, 1]) * H - 0.5
                        x0, y0 = math.floor(ix), math.floor(iy)
                        fx, fy = ix - x0, iy - y0
                        for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                                            (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
                            xx, yy = x0 + dx, y0 + dy
                            if 0 <= xx < W and 0 <= yy < H:
                                out[n, q, m] += float(attn[n, q, m, l, p]) * wgt * value[n, starts[l] + yy * W + xx, m]
    return out.reshape(N, Lq, M * D)


def test_grid_sample_definition_equals_the_tap_loop():
    g = torch.Generator().manual_seed(0)
    shapes = [(5, 7), (3, 4), (2, 2)]
    S = sum(h * w for h, w in shapes)
    value = torch.randn(2, S, 2, 4, generator=g, dtype=torch.float64)
    loc = torch.rand(2, 6, 2, 3, 2, 2, generator=g, dtype=torch.float64) * 1.6 - 0.3       # some points outside
    loc[0, 0, 0, 0, 0] = torch.tensor([0.0, 0.0])
    loc[0, 0, 0, 0, 1] = torch.tensor([1.0, 1.0])
    loc[0, 1, 0, 1, 0] = torch.tensor([0.5 / 4, 0.5 / 3])                                    # a cell centre
    attn = torch.rand(2, 6, 2, 6, generator=g, dtype=torch.float64).softmax(-1).view(2, 6, 2, 3, 2)
    torch.testing.assert_close(msda_torch(value, shapes, loc, attn), _loop(value, shapes, loc, attn),
                               rtol=1e-12, atol=1e-12)
