"""Time the DSG-FD streaming kernels alone (CUDA events, L2-exceeding inputs) for a few box coverages.

    python tools/kernel_sweep.py [--images 16] [--iters 20]
coverage 'none' = no teacher box (pure zero-fill of the gradient), 'full' = one box covering each image
(every byte is read), 'synth' = the bench's synthetic boxes.
"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dskd_b200 import _lib as L, synth  # noqa: E402


def build_args(inp, boxes, layout, with_grad=True):
    lib = L.load()
    dev = inp.hs_student.device
    st = L.stream_of(inp.hs_student)
    N, Q, Cc = inp.hs_student.shape
    levels, cells = L.levels_struct(inp.levels)
    lens = [len(b) for b in boxes]
    start = [0]
    for k in lens:
        start.append(start[-1] + k)
    meta = torch.tensor(start + [800, 1333] * N, dtype=torch.int32, device=dev)
    cat = torch.cat(boxes).float().contiguous() if sum(lens) else torch.zeros(0, 4, device=dev)
    P = cat.shape[0]
    owner = torch.empty(N, cells, dtype=torch.int32, device=dev)
    L.check(lib.dskd_raster_cells(L.RASTER_OWNER_EXCL, L.ptr(cat), L.ptr(meta[:N + 1]), None, None, L.ptr(meta[N + 1:]), N,
                                  max(lens + [0]), levels, len(inp.levels), cells, L.ptr(owner), st))
    rows = torch.softmax(torch.randn(max(P, 1), Cc, device=dev), 1).contiguous()
    energy = torch.zeros(max(P, 1), Cc, device=dev)
    a = L.DsgfdMseArgs()
    a.layout, a.num_levels, a.N, a.C = layout, len(inp.levels), N, Cc
    a.levels = levels
    a.cells_per_image = cells
    keep = []
    if layout == L.LAYOUT_NCHW:
        for l, (s, t) in enumerate(zip(inp.student_feats, inp.teacher_feats)):
            g = torch.empty_like(s) if with_grad else None
            keep.append(g)
            a.d_student[l], a.d_teacher[l] = s.data_ptr(), t.data_ptr()
            a.d_grad_student[l] = g.data_ptr() if with_grad else None
            a.scale[l] = 1.0 / N
    else:
        s, t = inp.memory()
        g = torch.empty_like(s) if with_grad else None
        keep += [s, t, g]
        a.d_student[0], a.d_teacher[0] = s.data_ptr(), t.data_ptr()
        a.d_grad_student[0] = g.data_ptr() if with_grad else None
        for l in range(len(inp.levels)):
            a.scale[l] = 1.0 / N
    a.d_owner, a.d_rows, a.d_energy, a.num_pairs = owner.data_ptr(), rows.data_ptr(), energy.data_ptr(), P
    keep += [owner, rows, energy, meta, cat]
    cov = float((owner >= 0).float().mean())
    return a, keep, cov


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=16)
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    lib = L.load()
    N = args.images
    inp = synth.make_distill_inputs(num_images=N, num_prev=40, seed=1234, device=dev)
    st = L.stream_of(inp.hs_student)
    full = [torch.tensor([[0., 0., 1333., 800.]], device=dev)] * N
    none = [torch.zeros(0, 4, device=dev)] * N
    per_img = 22223 * 256 * 4
    for layout, lname in ((L.LAYOUT_NCHW, 'nchw'), (L.LAYOUT_SNC, 'snc')):
        for cname, boxes in (('none', none), ('synth', inp.assignments['teacher_bboxes']), ('full', full)):
            a, keep, cov = build_args(inp, boxes, layout)
            for _ in range(3):
                L.check(lib.dskd_dsgfd_mse_fwd_bwd(a, st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                L.check(lib.dskd_dsgfd_mse_fwd_bwd(a, st))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            alg = 3 * per_img * N
            moved = (1 + 2 * cov) * per_img * N
            print(f'{lname:5s} {cname:6s} coverage {cov:5.2f}  {ms * 1e3:8.1f} us  algorithmic {alg / ms / 1e6:7.0f} GB/s  '
                  f'~moved {moved / ms / 1e6:7.0f} GB/s', flush=True)


if __name__ == '__main__':
    main()
