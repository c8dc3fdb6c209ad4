#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -q -x > gpurun_out/r2_tests17.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_tests17.log
cat > /tmp/deepbench.py <<P
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rgbx_semantic_segmentation_b200 import ops
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, N, K, tb) in [(19200, 320, 1280, False), (19200, 1280, 320, False), (4800, 512, 2048, False), (4800, 2048, 512, False), (76800, 128, 512, False), (153600, 512, 512, False), (19200, 1280, 320, True), (4800, 2048, 512, True), (19200, 320, 1280, True)]:
    a = torch.randn(M, K, device="cuda").to(bf); b = (torch.randn(K, N, device="cuda") if tb else torch.randn(N, K, device="cuda")).to(bf)
    out = torch.empty(M, N, device="cuda", dtype=bf); bias = None if tb else torch.randn(N, device="cuda")
    ts = []
    for i in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.mm(a, b, out, tb=tb, bias=bias, impl=2); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print("M=%d N=%d K=%d tb=%d: %.1f us  %.1f TFLOP/s" % (M, N, K, tb, t, 2.0 * M * N * K / t / 1e6), flush=True)
P
for dk in 512 0 320; do echo "== CMX_GEMM_DEEP_K=$dk"; CMX_GEMM_DEEP_K=$dk timeout 120 python /tmp/deepbench.py; done
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model17.log 2>&1
echo "model rc=$?"; tail -2 gpurun_out/r2_model17.log
for dk in 512 0 320 512 0; do
CMX_GEMM_DEEP_K=$dk timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench17_$dk.json 2> gpurun_out/r2_bench17_$dk.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench17_$dk.json"))
print("DEEP_K=$dk ms/step", d["ms_per_step"], "img/s", d["value"], "infer", d["inference"]["batch8"]["ms_per_forward"])
P
done
