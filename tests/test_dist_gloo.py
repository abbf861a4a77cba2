"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in dskd_b200/dist.py.

The CUDA kernels cannot run here, so each rank builds its LOCAL prototype table with the oracle (the checker),
the product's `allreduce_prototypes` combines them, and the result must equal the oracle on the concatenated
batch: prototypes, loss, and -- after the x world_size adjoint and DDP's gradient mean -- the gradient.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dskd_b200 import dist as dd
from dskd_b200 import synth

Q, C, L = 40, 32, 10
LEVELS = ((6, 10), (3, 5))


def _inputs(n_images):
    return synth.make_distill_inputs(num_images=n_images, num_prev=L, seed=99, levels=LEVELS, num_query=Q,
                                     channels=C, img_hw=(96, 160), k_range=(3, 8))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, ret):
    from oracle import bcdd as ob, losses as ol
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        full = _inputs(n_images)
        a = full.assignments
        crit = ol.MSELoss('mean', 1.0)
        # ---- single-process oracle on the concatenated batch
        hs_all = full.hs_student.reshape(-1, C).clone().requires_grad_(True)
        ht_all = full.hs_teacher.reshape(-1, C)
        ct, cs = ob.prototypes(hs_all, a['student_labels'], ht_all, a['teacher_keepid'], a['teacher_labels'],
                               a['prev_labels'])
        proto_ref = torch.stack([ct, cs]).detach().clone()
        loss_ref = ob.correlation_loss(ct, cs, L, crit)
        loss_ref.backward()

        # ---- this rank's shard
        b, e = dd.shard_bounds(n_images, rank, world)
        mine = dd.shard_assignments(a, Q, rank, world)
        hs = full.hs_student[b:e].reshape(-1, C).clone().requires_grad_(True)
        ht = full.hs_teacher[b:e].reshape(-1, C)
        assert mine['student_labels'].numel() == (e - b) * Q
        assert int(mine['teacher_keepid'].min()) >= 0 and int(mine['teacher_keepid'].max()) < (e - b) * Q
        lt, ls = ob.prototypes(hs, mine['student_labels'], ht, mine['teacher_keepid'], mine['teacher_labels'],
                               a['prev_labels'])
        local = torch.stack([lt, ls])
        table = local.detach().clone().contiguous()
        grad_scale, work = dd.allreduce_prototypes(table)
        assert work is None and grad_scale == float(world)
        torch.testing.assert_close(table, proto_ref, rtol=1e-5, atol=1e-5)
        # counts are small integers: exact
        assert torch.equal(table[..., -1], proto_ref[..., -1])

        # loss on the reduced table; gradient flows through the local contribution only (the all-reduce's adjoint)
        glob = local + (table - local.detach())
        loss = ob.correlation_loss(glob[0], glob[1], L, crit)
        torch.testing.assert_close(loss.detach(), loss_ref.detach(), rtol=1e-4, atol=1e-7)
        loss.backward()
        ddp_grad = hs.grad * grad_scale / world        # x world_size adjoint, then DDP's mean over ranks
        ref_grad = hs_all.grad[b * Q:e * Q]
        scale = float(ref_grad.abs().max())
        torch.testing.assert_close(ddp_grad, ref_grad, rtol=1e-3, atol=1e-6 * max(scale, 1e-30) + 1e-9)

        # async handle + timing helper
        t2 = local.detach().clone().contiguous()
        gs, work = dd.allreduce_prototypes(t2, async_op=True)
        work.wait()
        torch.testing.assert_close(t2, table)
        assert dd.max_over_ranks(float(rank + 1), 'cpu') == float(world)
        ret[rank] = 'ok'
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_images', [4, 5])
def test_prototype_allreduce_matches_concatenated_batch(n_images):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_images, ret), nprocs=world, join=True)
        assert dict(ret) == {0: 'ok', 1: 'ok'}


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = dd.shard_bounds(n, r, world)
                assert 0 <= b <= e <= n and (e - b) in (n // world, n // world + 1)
                seen += list(range(b, e))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        dd.shard_bounds(4, 2, 2)


def test_single_process_is_a_no_op():
    t = torch.arange(12, dtype=torch.float32).reshape(2, 2, 3)
    gs, work = dd.allreduce_prototypes(t)
    assert gs == 1.0 and work is None and torch.equal(t, torch.arange(12, dtype=torch.float32).reshape(2, 2, 3))
    assert dd.max_over_ranks(3.5, 'cpu') == 3.5
    full = _inputs(3)
    same = dd.shard_assignments(full.assignments, Q, 0, 1)
    assert torch.equal(same['teacher_keepid'], full.assignments['teacher_keepid'])
    assert torch.equal(same['student_labels'], full.assignments['student_labels'])
