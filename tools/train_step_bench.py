#!/usr/bin/env python
"""40+40 incremental training step (BASELINE.json configs[1]) on synthetic data.

    python tools/train_step_bench.py --images-per-gpu 2 --steps 5 --warmup 2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        tools/train_step_bench.py --steps 5 --warmup 2

Random-init GFL-Deformable-DETR R-50 student + frozen deep-copied teacher (dskd_b200/harness), 800x1333 synthetic
images, AdamW + grad-clip 0.1, fp32 (TF32 tensor-core math allowed for convolutions / matmuls, PyTorch's default for
conv).  The distillation path inside the step is the CUDA one: teacher keep-ids, batched Hungarian assignment, BCDD,
DSG-FD (decode_v1; --criterion kl is the shipped config).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images-per-gpu', type=int, default=2)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--criterion', default='kl', choices=['kl', 'mse'])
    ap.add_argument('--height', type=int, default=800)
    ap.add_argument('--width', type=int, default=1333)
    ap.add_argument('--backbone', default='resnet50')
    ap.add_argument('--channels-last', action='store_true', help='NHWC convolutions in backbone and neck')
    ap.add_argument('--no-graphs', action='store_true', help='eager launches for the detectors too')
    args = ap.parse_args()
    import torch.distributed as dist
    from dskd_b200.harness import bench_train_step

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    line = bench_train_step(dev, rank, world, dist, images_per_gpu=args.images_per_gpu, criterion=args.criterion,
                            steps=args.steps, warmup=args.warmup, height=args.height, width=args.width, backbone=args.backbone,
                            graphs=not args.no_graphs, channels_last=args.channels_last)
    if rank == 0:
        line.update({'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None})
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        from dskd_b200 import peer
        peer.close_all()            # unmap the peers' NVLink buffers before anybody leaves
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
