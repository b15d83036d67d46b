import torch
dev = torch.device('cuda:0')
x = torch.randn(76085, 768, device=dev)
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
nb = x.numel() * 4
us = t(lambda: x.sum()); print("sum   us %.1f  read GB/s %.0f" % (us, nb / us / 1e3))
us = t(lambda: x.sum(dim=0)); print("sum0  us %.1f  read GB/s %.0f" % (us, nb / us / 1e3))
us = t(lambda: y.copy_(x)); print("copy  us %.1f  r+w GB/s %.0f" % (us, 2 * nb / us / 1e3))
us = t(lambda: x.abs().max()); print("absmax us %.1f" % us)
us = t(lambda: y.zero_()); print("zero  us %.1f  write GB/s %.0f" % (us, nb / us / 1e3))
