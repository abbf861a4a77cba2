"""Throughput of the tcgen05 query x memory kernel (run on the GPU box): TFLOP/s from 2*K*C*S flop per image.

The call (gather + contraction + combine) is captured once in a CUDA graph and replayed, so the time is device
time without host gaps; a 256 MB write between replays flushes the 126 MB L2.
    python tools/qmem_perf.py                 # sweep (BASELINE.json configs[4]: queries 100-900, batch 2/16)
    python tools/qmem_perf.py one N K [mode [S]]  # a single point; mode 1 = single-CTA kernel, 2 = CTA-pair kernel
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dskd_b200 import qmem


def bench(N, K, S=22223, C=256, Q=None, iters=15, mode=None):
    if mode is None:
        os.environ.pop('DSKD_QMEM_MODE', None)
    else:
        os.environ['DSKD_QMEM_MODE'] = str(mode)
    dev = 'cuda:0'
    Q = Q or max(300, K)
    g = torch.Generator(device=dev).manual_seed(0)
    mem = torch.randn(S, N, C, device=dev, generator=g)
    hs = torch.randn(N, Q, C, device=dev, generator=g)
    keep = torch.cat([torch.randperm(Q, device=dev)[:K] + i * Q for i in range(N)])
    sc = torch.rand(N * K, device=dev)
    start = torch.arange(N + 1, device=dev, dtype=torch.int32) * K
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            qmem.qmem_cell_weights(mem, hs, keep, sc, start, K)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        w = qmem.qmem_cell_weights(mem, hs, keep, sc, start, K)
    times = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[len(times) // 2]
    flop = 2.0 * K * C * S * N
    tf = flop / ms / 1e9
    print(f'mode={mode or "auto"} N={N:2d} K={K:3d} S={S}: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s ({100 * tf / 811.2:5.1f}% of tf32 = '
          f'measured bf16 1622.3 / 2)  memory {S * N * C * 4 / ms / 1e6:7.1f} GB/s  {N / ms * 1e3:9.0f} img/s', flush=True)
    return ms


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == 'one':
        bench(int(sys.argv[2]), int(sys.argv[3]), mode=sys.argv[4] if len(sys.argv) > 4 else None,
              S=int(sys.argv[5]) if len(sys.argv) > 5 else 22223)
        sys.exit(0)
    for N in (2, 16):
        for K in (100, 160, 300, 600, 900):
            for mode in ((1, 2) if K > 100 else (1,)):
                bench(N, K, mode=mode)
