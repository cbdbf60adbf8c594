import torch, time
x = torch.randint(-30000, 30000, (288, 1800000), dtype=torch.int16, device="cuda")
v = x.view(torch.int32)
for fn, name, nbytes in ((lambda: v.sum(), "sum int32 view (read-only)", x.numel()*2),
                         (lambda: x.max(), "max int16 (read-only)", x.numel()*2),
                         (lambda: v.clone(), "clone (copy: read+write)", x.numel()*4)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)/20
    print(f"{name}: {ms*1e3:.1f} us -> {nbytes/ms/1e6:.0f} GB/s")
