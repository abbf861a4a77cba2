"""The NVLink peer-memory prototype exchange (dskd_b200/peer.py, csrc/peer.cu) against the NCCL all-reduce: world-size-2
(or more) processes on one box, eager calls with alternating slots, a CUDA-graph replay, and the BCDD module end to end.
Needs at least two GPUs: skipped on a one-GPU box (the N = 1 path has no exchange)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import dskd_b200
        from dskd_b200 import dist as dskd_dist, peer, synth
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        numel = 2 * 80 * 257
        # ---- eager calls: both slots, many rounds
        worst = 0.0
        for it in range(12):
            table = torch.randn(numel, device=dev, generator=g)
            ref = table.clone()
            dist.all_reduce(ref)
            scale, _ = dskd_dist.allreduce_prototypes(table)
            assert scale == float(world)
            worst = max(worst, float((table - ref).abs().max() / ref.abs().max()))
        used_peer = bool(peer._exchanges)
        # every rank holds the bit-identical sum (rank-ordered addition)
        mine = table.clone()
        other = table.clone()
        dist.broadcast(other, src=0)
        identical = bool(torch.equal(mine, other))
        # ---- captured in a CUDA graph, replayed with fresh inputs
        static = torch.zeros(numel, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            dskd_dist.allreduce_prototypes(static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            dskd_dist.allreduce_prototypes(static)
        graph_worst = 0.0
        for it in range(5):
            fresh = torch.randn(numel, device=dev, generator=g)
            ref = fresh.clone()
            dist.all_reduce(ref)
            static.copy_(fresh)
            graph.replay()
            torch.cuda.synchronize()
            graph_worst = max(graph_worst, float((static - ref).abs().max() / ref.abs().max()))
        # ---- the module: synced prototypes through the peer kernel == NCCL transport
        inp = synth.make_distill_inputs(num_images=2, num_prev=40, seed=7 + rank, device=dev)
        losses = []
        for transport in ('nvlink', 'nccl'):
            os.environ['DSKD_PROTO_TRANSPORT'] = transport
            peer._unavailable.clear()
            if transport == 'nccl':
                saved, peer._exchanges = peer._exchanges, {}
            hs = inp.hs_student.detach().clone().requires_grad_(True)
            mod = dskd_b200.BetweenClassDistanceLoss(sync_prototypes=True)
            loss = mod(None, None, (hs, inp.hs_teacher), inp.assignments)
            loss.backward()
            losses.append((float(loss.detach()), hs.grad.clone()))
            if transport == 'nccl':
                peer._exchanges = saved
        os.environ.pop('DSKD_PROTO_TRANSPORT')
        rel = abs(losses[0][0] - losses[1][0]) / max(abs(losses[1][0]), 1e-30)
        gerr = float((losses[0][1] - losses[1][1]).abs().max() / losses[1][1].abs().max().clamp(min=1e-30))
        results[rank] = dict(worst=worst, used_peer=used_peer, identical=identical, graph_worst=graph_worst, loss_rel=rel,
                             grad_err=gerr)
        torch.cuda.synchronize()
        dist.barrier()
        peer.close_all()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one box')
def test_peer_exchange_matches_nccl():
    import torch.multiprocessing as mp
    world = 2          # the exchange is generic in the world size (bench.py checks 8 ranks); two keep the test light
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        out = dict(results)
    assert len(out) == world
    for rank, r in out.items():
        assert r['used_peer'], f'rank {rank}: the NVLink path was not taken (CUDA IPC / peer access unavailable?)'
        assert r['worst'] < 1e-6 and r['graph_worst'] < 1e-6, r
        assert r['identical'], 'the sum must be bit-identical on every rank'
        assert r['loss_rel'] < 1e-5 and r['grad_err'] < 1e-4, r
