#!/usr/bin/env python
"""BASELINE.json configs[4] in full: tokens {5k, 10k, 22k (COCO), 40k} x queries {100, 300, 600, 900} x images per GPU
{1, 2, 4, 8, 16}, the distillation step (DSG-FD decode_v1 + BCDD, forward + backward) graph-replayed on one B200;
masked MSE at every point, the shipped KL criterion at 300 queries.  Uses bench.py's `measure_config` (inputs that fit the
L2 are timed with an L2 flush between replays).  One line per point: step time, images/s, the streaming kernel's
event-timed duration and its fraction of the measured copy peak.

    python tools/full_sweep.py > profiles/r2/full_sweep.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dskd_b200 import synth  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print(f'{"crit":4s} {"tokens":>6s} {"queries":>7s} {"images":>6s} {"step us":>9s} {"images/s":>10s} {"kernel us":>10s} '
          f'{"GB/s":>7s} {"frac":>5s}  l2', flush=True)
    for crit, queries in (('mse', (100, 300, 600, 900)), ('kl', (300,))):
        for tokens in (5000, 10000, bench.COCO_TOKENS, 40000):
            levels = synth.COCO_LEVELS if tokens == bench.COCO_TOKENS else synth.scaled_levels(tokens)
            for q in queries:
                for n in (1, 2, 4, 8, 16):
                    fl = flush if 2 * n * tokens * 256 * 4 < (160 << 20) else None
                    r = bench.measure_config(dev, 0, 1, None, n, 40, crit, 6, 3, levels, q, fl)
                    print(f'{crit:4s} {r["tokens"]:6d} {q:7d} {n:6d} {r["ms_per_step"] * 1e3:9.1f} {r["value"]:10.0f} '
                          f'{r["kernel_ms"] * 1e3:10.1f} {r["achieved_gbs"]:7.0f} {r["frac"]:5.2f}  {r["l2"]}', flush=True)


if __name__ == '__main__':
    main()
