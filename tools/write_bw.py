"""Reference point for write-only kernels: what does a plain device fill achieve on this GPU?"""
import torch
x = torch.empty(8 * 1024**3, dtype=torch.int8, device="cuda")
for _ in range(3):
    x.fill_(1)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    x.fill_(1)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print(f"fill_ 8 GiB: {ms:.3f} ms -> {x.numel() / ms / 1e6:.1f} GB/s write-only")
y = torch.empty_like(x)
for _ in range(3):
    y.copy_(x)
torch.cuda.synchronize()
ev[0].record()
for _ in range(10):
    y.copy_(x)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
print(f"copy_ 8 GiB: {ms:.3f} ms -> {2 * x.numel() / ms / 1e6:.1f} GB/s read+write")
