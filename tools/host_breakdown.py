"""Where the host time of one eager distillation step goes (run on the GPU box): wall clock per piece, device idle
between pieces (every piece is timed over 300 repetitions with the queue drained before and after)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dskd_b200  # noqa: E402
from dskd_b200 import synth  # noqa: E402

dev = torch.device('cuda:0')
inp = synth.make_distill_inputs(num_images=int(os.environ.get('IMAGES', '16')), num_prev=40, seed=1234, device=dev)
dsg = dskd_b200.DSGFeatureDistillLoss(criterion=os.environ.get('CRIT', 'mse'))
bcdd = dskd_b200.BetweenClassDistanceLoss()
feats = [f.requires_grad_(True) for f in inp.student_feats]
hs = inp.hs_student.requires_grad_(True)
q = (hs, inp.hs_teacher)
REPS = 300


def timed(name, fn, reps=REPS):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'{name:44s} host {1e6 * (t1 - t0) / reps:7.1f} us   with drain {1e6 * (t2 - t0) / reps:7.1f} us', flush=True)


def clear():
    for f in feats:
        f.grad = None
    hs.grad = None


def full():
    clear()
    loss = dsg(feats, inp.teacher_feats, q, inp.assignments) + bcdd(None, None, q, inp.assignments)
    loss.backward()


def fwd_only():
    with torch.no_grad():
        dsg(feats, inp.teacher_feats, q, inp.assignments) + bcdd(None, None, q, inp.assignments)


def dsg_fwd():
    dsg(feats, inp.teacher_feats, q, inp.assignments)


def dsg_fwd_nograd():
    with torch.no_grad():
        dsg(feats, inp.teacher_feats, q, inp.assignments)


def bcdd_fwd():
    bcdd(None, None, q, inp.assignments)


def dsg_fb():
    clear()
    dsg(feats, inp.teacher_feats, q, inp.assignments).backward()


def bcdd_fb():
    clear()
    bcdd(None, None, q, inp.assignments).backward()


timed('full step (2 modules + add + backward)', full)
timed('both forwards, no_grad', fwd_only)
timed('DSG-FD forward (autograd on)', dsg_fwd)
timed('DSG-FD forward (no_grad)', dsg_fwd_nograd)
timed('BCDD forward (autograd on)', bcdd_fwd)
timed('DSG-FD forward + backward', dsg_fb)
timed('BCDD forward + backward', bcdd_fb)
timed('clear grads', clear)
x = torch.zeros(1, device=dev, requires_grad=True)
timed('torch: (x * 2 + x * 3).sum().backward()', lambda: (x * 2 + x * 3).sum().backward())
timed('torch.empty(1000 bytes, cuda)', lambda: torch.empty(1000, dtype=torch.uint8, device=dev))
