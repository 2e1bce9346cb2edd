import torch
x = torch.empty(2 * 1024**3, dtype=torch.uint8, device="cuda")
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.fill_(7)); print(f"fill (write only): {x.numel()/ms/1e6:.0f} GB/s")
ms = t(lambda: x.zero_()); print(f"memset: {x.numel()/ms/1e6:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy: {2*x.numel()/ms/1e6:.0f} GB/s (read+write)")
ms = t(lambda: x.sum(dtype=torch.int64)); print(f"sum (read only): {x.numel()/ms/1e6:.0f} GB/s")
