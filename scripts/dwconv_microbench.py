"""Standalone timing of the three depthwise-3x3 kernels (forward GELU, backward pre-activation pass, data gradient) at the four
stage shapes of the bench workload (MiT-B2 480x640, batch 8, both branches in one grouped launch).  Each kernel is captured into
a CUDA graph (10 launches, buffers larger than L2 at stages 1-2) and the graph replay is timed with CUDA events.
ITERS=1 (for ncu) launches every kernel once, eagerly."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbx_semantic_segmentation_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
ITERS = int(os.environ.get("ITERS", "10"))
ONLY = os.environ.get("ONLY")
PEAK = 6554.2
shapes = [(8, 120, 160, 256), (8, 60, 80, 512), (8, 30, 40, 1280), (8, 15, 20, 2048)]
torch.manual_seed(0)
s1 = torch.cuda.Stream()
torch.cuda.set_stream(s1)
for si, (B, H, W, C) in enumerate(shapes):
    if ONLY is not None and str(si) not in ONLY.split(","):
        continue
    M = 2 * B * H * W
    x = torch.randn(M, C, device=dev).bfloat16()
    dy = torch.randn(M, C, device=dev).bfloat16()
    y = torch.empty_like(x)
    du = torch.empty_like(x)
    dx = torch.empty_like(x)
    par = torch.randn(2, 16 * C, device=dev) * 0.2          # per-group parameter block: w [C,9] | bias [C] | padding
    gpar = torch.zeros(2, 16 * C, device=dev)
    w, b = par[0, :9 * C].view(C, 9), par[0, 9 * C:10 * C]
    gw, gb = gpar[0, :9 * C].view(C, 9), gpar[0, 9 * C:10 * C]
    gs = 16 * C
    runs = [("fwd", lambda: ops.dwconv3x3_fwd(x, w, b, ops.ACT_GELU, y, B, H, W, groups=2, param_gs=gs), 2 * x.numel() * 2),
            ("bwd_pre", lambda: ops.dwconv3x3_bwd_pre(x, w, b, ops.ACT_GELU, dy, du, gw, gb, B, H, W, groups=2, param_gs=gs), 3 * x.numel() * 2),
            ("dgrad", lambda: ops.dwconv3x3_fwd(du, w, None, ops.ACT_NONE, dx, B, H, W, flip=True, groups=2, param_gs=gs), 2 * x.numel() * 2)]
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
        print("stage %d %-8s [2x%dx%dx%d, C=%d]: %7.1f us  %5.0f GB/s  frac %.3f" % (si + 1, name, B, H, W, C, best, nbytes / best * 1e-3,
                                                                                  nbytes / best * 1e-3 / PEAK), flush=True)
