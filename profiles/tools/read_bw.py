"""Read-only streaming bandwidth of this box (torch.sum over a 2 GiB float32 tensor) next to the copy figure."""
import torch
x = torch.empty(512 * 1024 * 1024, dtype=torch.float32, device="cuda").normal_()
y = torch.empty_like(x)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.sum())
print(f"torch.sum  2 GiB read : {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s")
ms = t(lambda: x.max())
print(f"torch.max  2 GiB read : {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x))
print(f"copy 2 GiB -> 2 GiB   : {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s (read+write)")
z = x[: 137 * 1024 * 1024]
ms = t(lambda: z.sum())
print(f"torch.sum  548 MB read: {ms:.3f} ms  {z.numel()*4/ms/1e6:.0f} GB/s")
