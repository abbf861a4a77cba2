"""Graph-replayed fwd+bwd time of DSGFeatureDistillLoss for every mask mode / layout at the bench shape (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dskd_b200
from dskd_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = 'cuda:0'
inp = synth.make_distill_inputs(num_images=N, num_prev=40, seed=1234, device=dev)
a = dict(inp.assignments)
a['teacher_scores'] = torch.rand(a['teacher_keepid'].numel(), device=dev)
s_mem, t_mem = inp.memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = [('decode_v1', 'mse', 'neck'), ('decode_v1', 'kl', 'neck'), ('decode_v1', 'mse', 'memory'), ('decode_v2', 'mse', 'neck'),
         ('sg_out', 'mse', 'memory'), ('fg_only', 'mse', 'memory'), ('fg_bk', 'mse', 'memory'), ('sg_out', 'kl', 'neck'),
         ('sg_out', 'kl', 'memory'), ('fg_only', 'kl', 'neck'), ('fg_only', 'kl', 'memory'), ('decode_v2', 'kl', 'neck'),
         ('qmem', 'mse', 'memory')]
for mode, crit, src in CASES:
    mod = dskd_b200.build_loss(dict(type='DSGFeatureDistillLoss', criterion=crit, mask_mode=mode, feature_source=src))
    if src == 'neck':
        sf = [f.clone().requires_grad_(True) for f in inp.student_feats]
        tf = inp.teacher_feats
        leaves = sf
    else:
        m = s_mem.clone().requires_grad_(True)
        sf, tf, leaves = (m, inp.spatial_shapes), (t_mem, inp.spatial_shapes), [m]
    hs = inp.hs_student.clone().requires_grad_(True)

    def step():
        for l in leaves:
            l.grad = None
        hs.grad = None
        loss = mod(sf, tf, (hs, inp.hs_teacher), a)
        loss.backward()
        return loss
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for l in leaves:
        l.grad = None
    hs.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    alg = (3 if crit == 'mse' else 2) * 22223 * 256 * 4 * N
    print(f'{mode:10s} {crit:3s} {src:6s}: {ms * 1e3:8.1f} us  {N / ms * 1e3:9.0f} img/s  algorithmic {alg / ms / 1e6:7.0f} GB/s  loss {float(loss.detach()):.5g}', flush=True)
