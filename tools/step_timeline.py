"""Kernel timeline of ONE replay of the captured distillation step (torch.profiler / CUPTI): start, duration, stream.
    python tools/step_timeline.py [mse|kl] [images]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

crit = sys.argv[1] if len(sys.argv) > 1 else 'mse'
images = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device('cuda:0')
torch.cuda.set_device(0)
torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
sb = bench.StepBench(dev, 0, 1, None, images, 40, crit)
runner, loss, policy = sb.capture()
for _ in range(5):
    runner.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        runner.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last replay only
n = len(evs) // 3
evs = evs[-n:]
t0 = evs[0].time_range.start
print(f'{policy}; {n} device activities per replay')
end_prev = t0
for e in evs:
    st, en = e.time_range.start - t0, e.time_range.end - t0
    print(f'{st:8.1f} -> {en:8.1f} us  ({en - st:6.1f})  {e.name[:90]}')
print(f'replay span {evs[-1].time_range.end - t0:.1f} us')
