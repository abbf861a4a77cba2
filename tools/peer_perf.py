#!/usr/bin/env python
"""Latency of the prototype exchange (164 KB table): the NVLink peer-memory kernel against the NCCL all-reduce, back to
back on one stream, CUDA events, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tools/peer_perf.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dskd_b200 import dist as dskd_dist, peer  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    table = torch.randn(2 * 80 * 257, device=dev)

    def timed(fn, iters=300):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    nccl_us = timed(lambda: dist.all_reduce(table))
    table.normal_()
    peer_us = timed(lambda: dskd_dist.allreduce_prototypes(table))
    used = bool(peer._exchanges)
    # the same inside CUDA graphs (what the bench step replays): 50 calls per graph
    def graphed(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                fn()
        us = timed(g.replay, iters=20) / 50
        del g
        return us
    table.normal_()
    nccl_graph_us = graphed(lambda: dist.all_reduce(table))
    table.normal_()
    peer_graph_us = graphed(lambda: dskd_dist.allreduce_prototypes(table))
    if rank == 0:
        print(f'world {world}: NCCL all-reduce {nccl_us:.1f} us eager / {nccl_graph_us:.1f} us in a graph; peer-memory kernel '
              f'{peer_us:.1f} us eager / {peer_graph_us:.1f} us in a graph (peer path used: {used})', flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    peer.close_all()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
