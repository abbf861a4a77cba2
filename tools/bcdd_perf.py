"""Device time of the BCDD chain (prototypes -> distances / loss / gradient / scatter) in a CUDA graph (run on the GPU box):
the module call, and the two C-ABI entry points on their own."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dskd_b200
from dskd_b200 import _lib as L, synth


def replay_us(fn, reps=50):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for num_prev in (40, 70):
    inp = synth.make_distill_inputs(num_images=16, num_prev=num_prev, seed=1234, device='cuda:0', levels=((4, 4),))
    mod = dskd_b200.build_loss(dict(type='BetweenClassDistanceLoss', reduction='mean'))
    hs = inp.hs_student.requires_grad_(True)

    def step():
        hs.grad = None
        loss = mod(None, None, (hs, inp.hs_teacher), inp.assignments)
        loss.backward()
        return loss
    chain = replay_us(step)
    lib = L.load()
    a = inp.assignments
    dev = hs.device
    C, nc = hs.shape[-1], 80
    hs2, ht2 = hs.detach().reshape(-1, C), inp.hs_teacher.reshape(-1, C)
    labels, keep, tl = a['student_labels'], a['teacher_keepid'], a['teacher_labels']
    prev = torch.zeros(nc, dtype=torch.uint8, device=dev)
    prev[:num_prev] = 1
    proto = torch.empty(2, nc, C + 1, device=dev)
    dist = torch.empty(2, num_prev, num_prev, device=dev)
    loss = torch.empty(1, device=dev)
    gp = torch.empty(nc, C + 1, device=dev)
    gh = torch.empty_like(hs2)

    def protos():
        L.check(lib.dskd_bcdd_prototypes(L.ptr(hs2), L.ptr(labels), hs2.shape[0], L.ptr(ht2), L.ptr(keep), L.ptr(tl),
                                         keep.numel(), L.ptr(prev), nc, C, L.ptr(proto), L.stream_of(hs2)))

    def tail():
        L.check(lib.dskd_bcdd_loss_and_grad(L.ptr(proto), nc, C, num_prev, L.REDUCTION_MEAN, 1.0, 1.0, L.ptr(labels),
                                            hs2.shape[0], L.ptr(prev), L.ptr(dist), L.ptr(loss), L.ptr(gp), L.ptr(gh),
                                            L.stream_of(hs2)))
    protos()
    print(f'L={num_prev}: BCDD fwd+bwd chain {chain:.1f} us per replay; dskd_bcdd_prototypes {replay_us(protos):.1f} us, '
          f'dskd_bcdd_loss_and_grad {replay_us(tail):.1f} us', flush=True)
