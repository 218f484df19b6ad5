"""How much of an in-place read-modify-write working set does the B200 L2 keep between kernels?"""
import torch
dev = torch.device('cuda:0')
for mb in (8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 512, 1024):
    n = mb * 1024 * 1024 // 8
    x = torch.zeros(n, dtype=torch.double, device=dev)
    for _ in range(5):
        x.add_(1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(20, 4096 // mb)
    e0.record()
    for _ in range(reps):
        x.add_(1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'{mb:5d} MB in-place add: {ms * 1e3:8.1f} us/pass  {2 * mb / 1024 / (ms * 1e-3):8.1f} GB/s (R+W)')
