#!/usr/bin/env python
"""Where the 40+40 training step spends its device time (torch.profiler, 2 steps after 2 warm-up steps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from torch.profiler import ProfilerActivity, profile
    from dskd_b200.harness import IncrementalTrainStep, make_student_teacher
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    crit = sys.argv[2] if len(sys.argv) > 2 else 'kl'
    dev = torch.device('cuda', 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    student, teacher = make_student_teacher(dev)
    student.train()
    trainer = IncrementalTrainStep(student, teacher, num_prev=40, criterion=crit)
    g = torch.Generator(device=dev).manual_seed(1234)
    H, W = 800, 1333
    img = torch.randn(n, 3, H, W, device=dev, generator=g)
    gt_b, gt_l = [], []
    for _ in range(n):
        k = 5
        x1 = torch.rand(k, device=dev, generator=g) * 0.7 * W
        y1 = torch.rand(k, device=dev, generator=g) * 0.7 * H
        gt_b.append(torch.stack([x1, y1, (x1 + 200).clamp(max=W), (y1 + 150).clamp(max=H)], 1))
        gt_l.append(torch.randint(40, 80, (k,), device=dev, generator=g))
    for _ in range(2):
        trainer.step(img, gt_b, gt_l)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            trainer.step(img, gt_b, gt_l)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=70))


if __name__ == '__main__':
    main()
