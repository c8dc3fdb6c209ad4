#!/bin/bash
# Round 2 call 9: GEMM epilogue with 8 warps (2 CTAs/SM) + leaner body: tests, microbench A/B, step A/B
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -q -x > gpurun_out/r2_ops_tests.log 2>&1
echo "ops rc=$?"; tail -3 gpurun_out/r2_ops_tests.log
for ew in 8 4; do
echo "== CMX_GEMM_EPI_WARPS=$ew"
CMX_GEMM_EPI_WARPS=$ew ONLY=8,9,10,11,12,13,14,15,w0,w1,w2,w3 ITERS=7 timeout 200 python scripts/gemm_microbench.py 2>&1 | tail -14
done
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model_tests.log 2>&1
echo "model rc=$?"; tail -3 gpurun_out/r2_model_tests.log
for ew in 8 4 8 4; do
CMX_GEMM_EPI_WARPS=$ew timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ew$ew.json 2> gpurun_out/r2_bench_ew$ew.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench_ew$ew.json"))
print("EW=$ew ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "infer", d["inference"]["batch8"]["ms_per_forward"])
P
done
