"""Standalone timing of the fused attention kernels at the four stage shapes of the bench workload (MiT-B2 480x640, batch 8, both
branches stacked): training forward (stores P), inference forward (no P, scaled O), stored-P backward core.  Each kernel is captured
into a CUDA graph (10 launches) and the replay is timed with CUDA events.  ITERS=1 launches every kernel once (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbx_semantic_segmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
ITERS = int(os.environ.get("ITERS", "10"))
ONLY = os.environ.get("ONLY")
PEAK = 6554.2
shapes = [(16, 19200, 300, 1), (16, 4800, 300, 2), (16, 1200, 300, 5), (16, 300, 300, 8)]
torch.manual_seed(0)
s1 = torch.cuda.Stream()
torch.cuda.set_stream(s1)
for si, (B, N, Nk, heads) in enumerate(shapes):
    if ONLY is not None and str(si) not in ONLY.split(","):
        continue
    C = 64 * heads
    q = torch.randn(B * N, C, device=dev).bfloat16()
    kv = torch.randn(B * Nk, 2 * C, device=dev).bfloat16()
    do = torch.randn(B * N, C, device=dev).bfloat16()
    o = torch.empty_like(q)
    dq = torch.empty_like(q)
    Np = (Nk + 7) // 8 * 8
    P = torch.empty(B * heads * N, Np, device=dev, dtype=torch.bfloat16)[:, :Nk]
    dS = torch.empty(B * heads * N, Np, device=dev, dtype=torch.bfloat16)[:, :Nk]
    scale = 0.125
    pbytes = B * heads * N * Nk * 2
    io = (q.numel() + o.numel() + kv.numel()) * 2
    runs = [("fwd_train", lambda: ops.attn_fwd(q, kv, o, B, N, Nk, heads, scale, p_out=P), io + pbytes),
            ("fwd_infer", lambda: ops.attn_fwd(q, kv, o, B, N, Nk, heads, scale), io),
            ("bwd", lambda: ops.attn_bwd(do, kv, P, dS, dq, B, N, Nk, heads, scale), io + 2 * pbytes)]
    for name, fn, nbytes in runs:
        fn()
        torch.cuda.synchronize()
        if ITERS == 1:
            continue
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s1):
            for _ in range(ITERS):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, 1e3 * e0.elapsed_time(e1) / ITERS)
        fl = 4 * B * heads * N * Nk * 64
        print("stage %d %-10s [B=%d N=%d Nkv=%d heads=%d]: %7.1f us  %5.0f GB/s (frac %.3f)  %6.1f TFLOP/s" % (
            si + 1, name, B, N, Nk, heads, best, nbytes / best * 1e-3, nbytes / best * 1e-3 / PEAK, fl / best * 1e-6), flush=True)
