"""Isolated timing of the tcgen05 GEMM on the bench's dominant shapes (CUDA events, L2 flushed between iterations)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbx_semantic_segmentation_b200 import ops  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [  # (M, N, K, out dtype, bias, residual fp32, tb)
    (128, 64, 64, bf, False, False, False), (128 * 148, 64, 64, bf, False, False, False),
    (128 * 148 * 2, 64, 64, bf, False, False, False), (128 * 148 * 4, 64, 64, bf, False, False, False),
    (128 * 148 * 8, 64, 64, bf, False, False, False), (128 * 148, 256, 64, bf, False, False, False),
    (128 * 148 * 4, 256, 64, bf, False, False, False), (128 * 148 * 8, 256, 64, bf, False, False, False),
    (153600, 64, 64, bf, True, False, False), (153600, 256, 64, bf, True, False, False),
    (153600, 64, 256, torch.float32, True, True, False), (153600, 64, 64, torch.float32, True, True, False),
    (38400, 512, 128, bf, True, False, False), (9600, 1280, 320, bf, True, False, False),
    (153600, 256, 64, bf, False, False, True), (153600, 512, 512, bf, False, False, False)]
iters = int(os.environ.get("ITERS", "5"))
only = os.environ.get("ONLY")
x8 = torch.zeros(8, device=dev)
y8 = torch.zeros(8, device=dev, dtype=bf)
ts = []
for i in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.cast_f32_bf16(x8, y8); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("event floor with a trivial kernel: %.1f us" % sorted(ts)[3])
for idx, (M, N, K, odt, hb, hr, tb) in enumerate(SHAPES):
    if only is not None and str(idx) not in only.split(","):
        continue
    a = torch.randn(M, K, device=dev).to(bf)
    b = (torch.randn(K, N, device=dev) if tb else torch.randn(N, K, device=dev)).to(bf)
    out = torch.empty(M, N, device=dev, dtype=odt)
    bias = torch.randn(N, device=dev) if hb else None
    res = torch.randn(M, N, device=dev) if hr else None
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mm(a, b, out, tb=tb, bias=bias, residual=res, impl=2)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    nbytes = a.numel() * 2 + b.numel() * 2 + out.numel() * out.element_size() + (res.numel() * 4 if hr else 0)
    t = sorted(ts)[len(ts) // 2]
    print("%d: M=%d N=%d K=%d out=%s bias=%d res=%d tb=%d : %.1f us  %.0f GB/s  %.1f TFLOP/s" % (
        idx, M, N, K, str(odt)[6:], hb, hr, tb, t, nbytes / t / 1e3, 2.0 * M * N * K / t / 1e6), flush=True)

# weight-gradient shapes: dW[N_out, N_in] += dY[tok, N_out]^T X[tok, N_in]  (MN-major operands, split-K fp32 atomics)
WGRAD = [(64, 64, 153600), (256, 64, 153600), (128, 128, 38400), (512, 128, 38400), (320, 320, 9600), (1280, 320, 9600),
         (512, 512, 2400), (512, 512, 153600)]
for idx, (M, N, K) in enumerate(WGRAD):
    if only is not None and ("w%d" % idx) not in only.split(","):
        continue
    dy = torch.randn(K, M, device=dev).to(bf)
    x = torch.randn(K, N, device=dev).to(bf)
    out = torch.zeros(M, N, device=dev)
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mm(dy, x, out, ta=True, tb=True, accumulate=True)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    nbytes = dy.numel() * 2 + x.numel() * 2 + out.numel() * 8
    t = sorted(ts)[len(ts) // 2]
    print("w%d: dW[%d,%d] over %d tokens : %.1f us  %.0f GB/s  %.1f TFLOP/s" % (idx, M, N, K, t, nbytes / t / 1e3,
                                                                             2.0 * M * N * K / t / 1e6), flush=True)
