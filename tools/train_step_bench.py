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
    args = ap.parse_args()
    import torch.distributed as dist
    from dskd_b200.harness import IncrementalTrainStep, make_student_teacher

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True

    student, teacher = make_student_teacher(dev, backbone=args.backbone)
    student.train()
    if world > 1:
        student = torch.nn.parallel.DistributedDataParallel(student, device_ids=[local_rank], broadcast_buffers=False,
                                                            find_unused_parameters=True)   # train_increment.py:301-303
    trainer = IncrementalTrainStep(student, teacher, num_prev=40, criterion=args.criterion, sync_prototypes=world > 1)
    N, H, W = args.images_per_gpu, args.height, args.width
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    img = torch.randn(N, 3, H, W, device=dev, generator=g)
    gt_b, gt_l = [], []
    for _ in range(N):
        k = int(torch.randint(1, 11, (1,), device=dev, generator=g))
        x1 = torch.rand(k, device=dev, generator=g) * 0.7 * W
        y1 = torch.rand(k, device=dev, generator=g) * 0.7 * H
        bw = 8 + torch.rand(k, device=dev, generator=g) * (0.3 * W - 8)
        bh = 8 + torch.rand(k, device=dev, generator=g) * (0.3 * H - 8)
        gt_b.append(torch.stack([x1, y1, (x1 + bw).clamp(max=W), (y1 + bh).clamp(max=H)], 1))
        gt_l.append(torch.randint(40, 80, (k,), device=dev, generator=g))

    out = None
    for _ in range(max(args.warmup, 1)):
        out = trainer.step(img, gt_b, gt_l)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = trainer.step(img, gt_b, gt_l)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        line = {'metric': 'incremental_train_step_images_per_s', 'value': world * N * args.steps / (ms * 1e-3),
                'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 1),
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'dtype': 'f32 (tf32 matmul/conv)',
                'data': 'synthetic', 'vs_baseline': None,
                'config': {'workload': 'coco_40+40_incremental_train_step', 'images_per_gpu': N, 'image': [H, W],
                           'backbone': args.backbone, 'criterion': args.criterion, 'queries': 300, 'decoder_layers': 6,
                           'parallelism': f'dp{world}'},
                'losses': {k: float(v) for k, v in out.items()},
                'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
