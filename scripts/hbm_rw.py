"""HBM read-only / write-only / copy bandwidth probe (torch kernels, CUDA events) - context for the roofline numbers."""
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
xf = x.view(torch.float32)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
gb = (1 << 30) / 1e9
print("write-only (zero_):  %.0f GB/s" % (gb / (t(lambda: x.zero_()) * 1e-3)))
print("read-only  (sum):    %.0f GB/s" % (gb / (t(lambda: xf.sum()) * 1e-3)))
print("copy (read+write):   %.0f GB/s" % (2 * gb / (t(lambda: y.copy_(x)) * 1e-3)))
