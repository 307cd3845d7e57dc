import torch, time
n = 32163584
h = torch.zeros(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty_like(d)
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2d", lambda: d2.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(name, "%.3f ms  %.1f GB/s" % (ms, n / ms / 1e6))
