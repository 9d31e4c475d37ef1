import torch, time
x = torch.empty(6 * 1024**3 // 8, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for name, fn, nbytes in (("fill (write only)", lambda: x.fill_(1.0), x.numel() * 8), ("copy (read+write)", lambda: y.copy_(x), 2 * x.numel() * 8)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print(name, round(best, 1), "GB/s")
