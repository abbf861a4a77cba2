"""cProfile of the host side of one distillation step (run on the GPU box)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dskd_b200  # noqa: E402
from dskd_b200 import synth  # noqa: E402

dev = torch.device('cuda:0')
inp = synth.make_distill_inputs(num_images=16, num_prev=40, seed=1234, device=dev)
dsg = dskd_b200.DSGFeatureDistillLoss(criterion='mse')
bcdd = dskd_b200.BetweenClassDistanceLoss()
feats = [f.requires_grad_(True) for f in inp.student_feats]
hs = inp.hs_student.requires_grad_(True)


def step():
    for f in feats:
        f.grad = None
    hs.grad = None
    loss = dsg(feats, inp.teacher_feats, (hs, inp.hs_teacher), inp.assignments) + \
        bcdd(None, None, (hs, inp.hs_teacher), inp.assignments)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host issue time {1e6 * (t1 - t0) / 200:.1f} us/step, with final sync {1e6 * (t2 - t0) / 200:.1f} us/step')
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumtime').print_stats(45)
