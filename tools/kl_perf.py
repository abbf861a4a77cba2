"""Time the DSG-FD KL-over-H streaming kernel alone (CUDA events around `dskd_dsgfd_kl_fwd_bwd`, inputs larger than L2).

    python tools/kl_perf.py [--images 16] [--iters 20] [--tune "2,4,16,192,0,8,1;2,4,8,192,0,8,2;..."]
coverage 'synth' = the bench's synthetic boxes, 'full' = one box covering each image (every byte is read).
--tune runs the box-mask case once per "channels_per_pass,rows_per_block,channels_per_cta,pool,dbg,warps,row_parts"
setting (DSKD_KL_TUNE; rows_per_block is 4 for the gradient kernel and 5 for the forward-only one: the forward timing of
a setting uses 5).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dskd_b200 import _lib as L, synth  # noqa: E402


def build_args(inp, boxes, cell=False, grad=True, layout=L.LAYOUT_NCHW):
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
    cat = torch.cat(boxes).float().contiguous()
    P = cat.shape[0]
    a = L.DsgfdKlArgs()
    a.num_levels, a.N, a.C = len(inp.levels), N, Cc
    a.levels = levels
    a.cells_per_image = cells
    a.temperature = 2.0
    if hasattr(a, 'layout'):
        a.layout = layout
    keep = [meta, cat]
    if layout == L.LAYOUT_NCHW:
        for l, (s, t) in enumerate(zip(inp.student_feats, inp.teacher_feats)):
            a.d_student[l], a.d_teacher[l] = s.data_ptr(), t.data_ptr()
    else:
        s, t = inp.memory()
        keep += [s, t]
        a.d_student[0], a.d_teacher[0] = s.data_ptr(), t.data_ptr()
    for l in range(len(inp.levels)):
        a.scale[l] = 1.0 / N
    if cell:
        w = torch.empty(N, cells, dtype=torch.float32, device=dev)
        L.check(lib.dskd_raster_cells(L.RASTER_AREA_INCL, L.ptr(cat), L.ptr(meta[:N + 1]), None, None, L.ptr(meta[N + 1:]),
                                      N, max(lens), levels, len(inp.levels), cells, L.ptr(w), st))
        a.d_cell_weight = w.data_ptr()
        cov = float((w != 0).float().mean())
        keep.append(w)
    else:
        owner = torch.empty(N, cells, dtype=torch.int32, device=dev)
        L.check(lib.dskd_raster_cells(L.RASTER_OWNER_EXCL, L.ptr(cat), L.ptr(meta[:N + 1]), None, None, L.ptr(meta[N + 1:]),
                                      N, max(lens), levels, len(inp.levels), cells, L.ptr(owner), st))
        rows = torch.softmax(torch.randn(P, Cc, device=dev).abs(), 1).contiguous()
        grad_rows = torch.zeros(P, Cc, device=dev)
        a.d_owner, a.d_rows, a.num_pairs = owner.data_ptr(), rows.data_ptr(), P
        a.d_grad_rows = grad_rows.data_ptr() if grad else None
        cov = float((owner >= 0).float().mean())
        keep += [owner, rows, grad_rows]
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    a.d_loss = loss.data_ptr()
    nbytes = lib.dskd_dsgfd_kl_workspace_bytes(N, len(inp.levels), levels, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    a.d_workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    keep += [loss, ws]
    return a, keep, cov, loss


def time_call(lib, a, st, iters):
    for _ in range(3):
        L.check(lib.dskd_dsgfd_kl_fwd_bwd(a, st), 'dskd_dsgfd_kl_fwd_bwd')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.check(lib.dskd_dsgfd_kl_fwd_bwd(a, st), 'dskd_dsgfd_kl_fwd_bwd')
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=16)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--tune', default='')
    ap.add_argument('--tune-only', action='store_true')
    ap.add_argument('--one', default='', help='run a single case "box|cell|snc_box|snc_cell" once per iteration (for ncu)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    lib = L.load()
    N = args.images
    inp = synth.make_distill_inputs(num_images=N, num_prev=40, seed=1234, device=dev)
    st = L.stream_of(inp.hs_student)
    full = [torch.tensor([[0., 0., 1333., 800.]], device=dev)] * N
    alg = 2 * 22223 * 256 * 4 * N
    has_layout = os.environ.get('DSKD_KL_SNC', '0') == '1'
    cases = [('nchw box+grad', dict()), ('nchw box fwd', dict(grad=False)), ('nchw cell', dict(cell=True))]
    if has_layout:
        cases += [('snc cell', dict(cell=True, layout=L.LAYOUT_SNC))]
    if args.one:
        sel = {'box': cases[0], 'cell': cases[2]}
        if has_layout:
            sel.update({'snc_cell': cases[3]})
        name, kw = sel[args.one]
        a, keep, cov, loss = build_args(inp, inp.assignments['teacher_bboxes'], **kw)
        ms = time_call(lib, a, st, args.iters)
        print(f'{name}: {ms * 1e3:.1f} us  algorithmic {alg / ms / 1e6:.0f} GB/s')
        return
    for name, kw in ([] if args.tune_only else cases):
        for cname, boxes in (('synth', inp.assignments['teacher_bboxes']), ('full', full)):
            a, keep, cov, loss = build_args(inp, boxes, **kw)
            ms = time_call(lib, a, st, args.iters)
            print(f'{name:14s} {cname:6s} coverage {cov:5.2f}  {ms * 1e3:8.1f} us  algorithmic {alg / ms / 1e6:7.0f} GB/s '
                  f'({alg / ms / 1e6 / 6516.7:.2f} of measured peak)', flush=True)
    if args.tune:
        a, keep, cov, loss = build_args(inp, inp.assignments['teacher_bboxes'])
        af, keepf, _, _ = build_args(inp, inp.assignments['teacher_bboxes'], grad=False)
        for setting in args.tune.split(';'):
            os.environ['DSKD_KL_TUNE'] = setting
            ms = time_call(lib, a, st, args.iters)
            f = setting.split(',')
            os.environ['DSKD_KL_TUNE'] = ','.join([f[0], '5'] + f[2:])
            msf = time_call(lib, af, st, args.iters)
            print(f'tune {setting:20s}: grad {ms * 1e3:8.1f} us {alg / ms / 1e6:7.0f} GB/s   fwd {msf * 1e3:8.1f} us '
                  f'{alg / msf / 1e6:7.0f} GB/s', flush=True)
        os.environ.pop('DSKD_KL_TUNE', None)


if __name__ == '__main__':
    main()
