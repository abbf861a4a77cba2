import torch, time
torch.backends.cuda.matmul.allow_tf32 = True
for dt, n in ((torch.float32, 8192), (torch.bfloat16, 8192)):
    a = torch.randn(n, n, device='cuda', dtype=dt); b = torch.randn(n, n, device='cuda', dtype=dt)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(dt, n, f'{ms:.3f} ms  {2 * n**3 / ms / 1e9:.1f} TFLOP/s')
