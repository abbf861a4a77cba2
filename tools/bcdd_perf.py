"""Device time of the BCDD chain (prototypes -> distances/loss/gradient -> scatter) in a CUDA graph (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dskd_b200
from dskd_b200 import synth

for L in (40, 70):
    inp = synth.make_distill_inputs(num_images=16, num_prev=L, seed=1234, device='cuda:0', levels=((4, 4),))
    mod = dskd_b200.build_loss(dict(type='BetweenClassDistanceLoss', reduction='mean'))
    hs = inp.hs_student.requires_grad_(True)

    def step():
        hs.grad = None
        loss = mod(None, None, (hs, inp.hs_teacher), inp.assignments)
        loss.backward()
        return loss
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    hs.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'L={L}: BCDD fwd+bwd chain {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per replay')
