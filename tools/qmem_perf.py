"""Throughput of the tcgen05 query x memory kernel (run on the GPU box): TFLOP/s from 2*K*C*S flop per image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dskd_b200 import qmem


def bench(N, K, S=22223, C=256, Q=None, iters=20):
    dev = 'cuda:0'
    Q = Q or max(300, K)
    g = torch.Generator(device=dev).manual_seed(0)
    mem = torch.randn(S, N, C, device=dev, generator=g)
    hs = torch.randn(N, Q, C, device=dev, generator=g)
    keep = torch.cat([torch.randperm(Q, device=dev)[:K] + i * Q for i in range(N)])
    sc = torch.rand(N * K, device=dev)
    start = torch.arange(N + 1, device=dev, dtype=torch.int32) * K
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        qmem.qmem_cell_weights(mem, hs, keep, sc, start, K)
    times = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        qmem.qmem_cell_weights(mem, hs, keep, sc, start, K)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[len(times) // 2]
    flop = 2.0 * K * C * S * N
    print(f'N={N:2d} K={K:3d} S={S} : {ms*1e3:8.1f} us  {flop/ms/1e9:7.1f} TFLOP/s  mem {S*N*C*4/ms/1e6:7.1f} GB/s '
          f'({N/ms*1e3:.0f} img/s)', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'one':
        bench(int(sys.argv[2]), int(sys.argv[3]), iters=3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'exp':
        for dbg in ('0', '1', '2', '3'):
            os.environ['DSKD_QMEM_DEBUG'] = dbg
            print('debug', dbg)
            bench(16, 160)
            bench(1, 160, S=22223 * 16)
            bench(16, 96)
        sys.exit(0)
    for N in (2, 16):
        for K in (100, 160, 300, 600, 900):
            bench(N, K)
