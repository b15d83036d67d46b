import sys, torch
sys.path.insert(0, '.')
from elimrec_b200 import ops
dev = torch.device('cuda:0')
I = 76085
dims = {'v': 128, 'a': 128, 't': 768}
X = {m: torch.randn(I, d, device=dev) for m, d in dims.items()}
W = {m: torch.randn(64, d, device=dev) for m, d in dims.items()}
b = {m: torch.randn(64, device=dev) for m in dims}
Y = torch.empty(I, 256, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
probs = [(X[m], W[m], b[m], Y, 64 * (j + 1)) for j, m in enumerate('vat')]
print("multi (v,a,t) us", t(lambda: ops.linear_tf32_fwd_multi(probs)))
print("t only us", t(lambda: ops.linear_tf32_fwd_multi(probs[2:])), " -> GB/s", I * 768 * 4 / 1e3 / t(lambda: ops.linear_tf32_fwd_multi(probs[2:])))
print("v only us", t(lambda: ops.linear_tf32_fwd_multi(probs[:1])))
