"""Per-shape timing of the bench's GEMMs (MiT-B2, 480x640, batch 8, one modality branch) for tile-policy tuning.
10 back-to-back launches replayed as one CUDA graph between two CUDA events (small operands stay L2-resident, as they are in the real step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbx_semantic_segmentation_b200 import ops  # noqa: E402

dev, bf, f32 = "cuda", torch.bfloat16, torch.float32
DIMS, TOK, HID = (64, 128, 320, 512), (153600, 38400, 9600, 2400), (512, 1024, 1280, 2048)
shapes = []  # (kind, M, N, K, out dtype)
for C, M, Hd in zip(DIMS, TOK, HID):
    shapes += [("fwd", M, C, C, bf), ("fwd", M, Hd, C, bf), ("fwd", M, C, Hd, f32), ("fwd", 2400, 2 * C, C, bf),
               ("dgrad", M, C, C, bf), ("dgrad", M, C, Hd, bf), ("dgrad", M, Hd, C, bf),
               ("wgrad", C, C, M, f32), ("wgrad", Hd, C, M, f32), ("wgrad", C, Hd, M, f32),
               ("fwd", M, 512, C, bf), ("fwd", M, 512, 512, bf), ("dgrad", M, 512, 512, bf), ("wgrad", 512, 512, M, f32)]
only = os.environ.get("KIND")
for kind, M, N, K, odt in shapes:
    if only and kind != only:
        continue
    if kind == "wgrad":
        a = torch.randn(K, M, device=dev).to(bf); b = torch.randn(K, N, device=dev).to(bf)
        out = torch.zeros(M, N, device=dev)
        run = lambda: ops.mm(a, b, out, ta=True, tb=True, accumulate=True)  # noqa: E731
    elif kind == "dgrad":
        a = torch.randn(M, K, device=dev).to(bf); b = torch.randn(K, N, device=dev).to(bf)
        out = torch.empty(M, N, device=dev, dtype=odt)
        run = lambda: ops.mm(a, b, out, tb=True)  # noqa: E731
    else:
        a = torch.randn(M, K, device=dev).to(bf); b = torch.randn(N, K, device=dev).to(bf)
        out = torch.empty(M, N, device=dev, dtype=odt)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev) if odt == f32 else None
        run = lambda: ops.mm(a, b, out, bias=bias, residual=res)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):   # graph replay: no host launch cost, like the real step
        for _ in range(10):
            run()
    g.replay()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 100)
    print("%-5s M=%6d N=%5d K=%6d : %6.1f us" % (kind, M, N, K, sorted(ts)[1]), flush=True)
