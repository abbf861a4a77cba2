#!/usr/bin/env python
"""Device time of the multi-scale deformable attention kernels at the detector's sizes (encoder: every token is a query;
decoder: 300 queries), CUDA events around a graph replay of forward / forward + backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dskd_b200.harness.msda import ms_deform_attn  # noqa: E402


def run(N, Lq_kind):
    dev = 'cuda:0'
    shapes = [(100, 167), (50, 84), (25, 42), (13, 21)]
    S = sum(h * w for h, w in shapes)
    Lq = S if Lq_kind == 'encoder' else 300
    g = torch.Generator(device=dev).manual_seed(0)
    value = torch.randn(N, S, 8, 32, device=dev, generator=g, requires_grad=True)
    # encoder: the query's own cell centre (raster order, as the transformer's reference points) plus a few cells of
    # offset; decoder: learned reference points, anywhere
    if Lq_kind == 'encoder':
        cs = []
        for h, w in shapes:
            ys, xs = torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h, (torch.arange(w, device=dev) + 0.5) / w,
                                    indexing='ij')
            cs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
        base = torch.cat(cs)[None, :, None, None, None, :].expand(N, Lq, 1, 1, 1, 2)
    else:
        base = torch.rand(N, Lq, 1, 1, 1, 2, device=dev, generator=g)
    loc = (base + 0.02 * torch.randn(N, Lq, 8, 4, 4, 2, device=dev, generator=g)).requires_grad_(True)
    attn = torch.rand(N, Lq, 8, 16, device=dev, generator=g).softmax(-1).view(N, Lq, 8, 4, 4).requires_grad_(True)
    go = torch.randn(N, Lq, 256, device=dev, generator=g)

    def fwd():
        return ms_deform_attn(value, shapes, loc, attn)

    def both():
        out = ms_deform_attn(value, shapes, loc, attn)
        out.backward(go)

    res = {}
    for name, fn in (('fwd', fwd), ('fwd+bwd', both)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[name] = ts[len(ts) // 2]
    taps = N * Lq * 8 * 16 * 4 * 128
    print(f'N={N} {Lq_kind:8s} Lq={Lq:6d}: fwd {res["fwd"] * 1e3:8.1f} us ({taps / res["fwd"] / 1e6:7.0f} GB/s of taps)  '
          f'fwd+bwd {res["fwd+bwd"] * 1e3:8.1f} us', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'one':        # a single encoder-sized point (ncu target)
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 2, 'encoder')
        sys.exit(0)
    for N in (2, 4):
        for kind in ('encoder', 'decoder'):
            run(N, kind)
