"""Multi-GPU host logic of the distillation hot path.

The reference is data-parallel only (mmcv DDP, tools/train_increment.py:299-304): every rank sees its own
images, DSG-FD and the Hungarian assignment are independent per image, and the reference's BCDD uses the
rank-local prototypes (gfl_deformable_detr_head_il.py:531-551).  The only cross-rank state this package
adds is the optional global prototype table: ONE all-reduce(sum) of the [2, num_classes, C+1] fp32 sums +
counts between `dskd_bcdd_prototypes` and `dskd_bcdd_loss_and_grad`.

The CPU-tensor path touches nothing but torch.distributed, so the same code runs under `gloo` (tests/test_dist_gloo.py,
world_size 2); CUDA tables on one NVLink box go through the library's own peer-memory kernel (dskd_b200/peer.py), anything
else through the NCCL all-reduce.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def shard_bounds(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `num_items` images owned by `rank`: the first `num_items % world_size`
    ranks hold one extra image, like torch's DistributedSampler without padding."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f'bad rank {rank} / world_size {world_size}')
    base, extra = divmod(int(num_items), world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_assignments(assignments: dict, num_query: int, rank: int, world_size: int) -> dict:
    """The slice of an `assignments` dict (dskd_b200.losses) that belongs to `rank`'s images: per-image lists are
    cut, `student_labels` [N*Q] is cut at image boundaries, and `teacher_keepid` (= q + Q*i, flattened over the
    GLOBAL batch, deformable_detr_il.py:151) is re-based to the local image index."""
    n = len(assignments['teacher_bboxes'])
    b, e = shard_bounds(n, rank, world_size)
    counts = [int(x.shape[0]) for x in assignments['teacher_bboxes']]
    p0, p1 = sum(counts[:b]), sum(counts[:e])
    out = dict(assignments)
    out['teacher_bboxes'] = list(assignments['teacher_bboxes'][b:e])
    if assignments.get('gt_bboxes') is not None:
        out['gt_bboxes'] = list(assignments['gt_bboxes'][b:e])
    out['img_shapes'] = assignments['img_shapes'][b:e]
    out['student_labels'] = assignments['student_labels'][b * num_query:e * num_query]
    out['teacher_keepid'] = assignments['teacher_keepid'][p0:p1] - b * num_query
    out['teacher_labels'] = assignments['teacher_labels'][p0:p1]
    return out


def allreduce_prototypes(proto: torch.Tensor, group=None, async_op: bool = False):
    """Sum the per-class prototype sums and counts over ranks, in place.

    proto: [2, num_classes, C+1] fp32 (0 = teacher, 1 = student; last column = count), as written by
    `dskd_bcdd_prototypes`.  Returns (grad_scale, work): grad_scale = world size -- every rank ends up with the
    same loss, so the adjoint of the all-reduce is a local x world_size, which DDP's gradient mean cancels;
    `work` is the async handle (None when not distributed or async_op is False)."""
    if not is_distributed(group):
        return 1.0, None
    if proto.dtype != torch.float32 or not proto.is_contiguous():
        raise ValueError('prototype table must be contiguous fp32')
    if proto.is_cuda and not async_op:
        # one kernel over NVLink peer memory instead of an NCCL launch (dskd_b200/peer.py); None when the ranks do not all
        # sit on one host with peer access -- the same decision on every rank
        from . import peer
        exchange = peer.exchange_for(proto, group)
        if exchange is not None:
            exchange.allreduce(proto)
            return float(dist.get_world_size(group)), None
    work = dist.all_reduce(proto, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return float(dist.get_world_size(group)), (work if async_op else None)


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing helper for bench.py: device-measured milliseconds, max over ranks."""
    if not is_distributed(group):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
