"""Dump the per-role clock64 timeline of CTA 0 of the persistent tcgen05 GEMM (debug)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgbx_semantic_segmentation_b200 import ops, _lib
bf = torch.bfloat16
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
a = torch.randn(M, K, device="cuda").to(bf); b = torch.randn(N, K, device="cuda").to(bf)
out = torch.empty(M, N, device="cuda", dtype=bf)
for _ in range(3):
    ops.mm(a, b, out, impl=2)
tr = torch.zeros(5 * 64 * 4, dtype=torch.int64, device="cuda")
_lib.load().cmx_debug_set_gemm_trace(tr.data_ptr())
ops.mm(a, b, out, impl=2)
torch.cuda.synchronize()
_lib.load().cmx_debug_set_gemm_trace(None)
t = tr.view(5, 64, 4).cpu()
t0 = int(t[t > 0].min())
ntile = (M + 127) // 128 * ((N + 63) // 64 if N <= 64 else 1)
n = min(12, -(-ntile // 148))
print("producer: slot-free time per tile:", [int(t[0, i, 0]) - t0 for i in range(n)])
for i in range(n):
    m = [int(v) - t0 for v in t[1, i]]
    e = [int(v) - t0 for v in t[2, i]]
    print("tile %2d  MMA: start %6d tempty-ok %6d mma-issued %6d committed %6d | EPI: top %6d bar-ok %6d tfull-ok %6d done %6d" % (i, *m, *e))

print("epilogue detail (warp 4, LAST unit of each tile): wait_read begin/end, chunk0 done, chunk1 done | fence+syncwarp done, store issued")
for i in range(n):
    d = [int(v) - t0 for v in t[3, i]] + [int(v) - t0 for v in t[4, i][:2]]
    print("tile %2d  %s" % (i, d))
