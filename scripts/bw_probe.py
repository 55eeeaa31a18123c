import torch, time
dev="cuda:0"
n = 1<<30  # 1 Gi floats = 4 GB
x = torch.empty(n, device=dev, dtype=torch.float32)
y = torch.empty(n, device=dev, dtype=torch.float32)
def t(fn, bytes_, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print("%-28s %8.3f ms  %7.1f GB/s" % (name, ms, bytes_/ms/1e6))
t(lambda: x.fill_(1.0), 4*n, "fill_ (pure write)")
t(lambda: x.zero_(), 4*n, "zero_ (memset)")
t(lambda: y.copy_(x), 8*n, "copy_ (read+write)")
t(lambda: x.sum(), 4*n, "sum (pure read)")
h = x.view(torch.int8)[:n]  # smaller
t(lambda: torch.add(x, 1.0, out=y), 8*n, "add out (r+w)")
xs = x[:n//2]; ys=y[:n//2]; z = torch.empty(n//2, device=dev)
t(lambda: torch.add(xs, ys, out=z), 12*(n//2), "add 2r+1w")
