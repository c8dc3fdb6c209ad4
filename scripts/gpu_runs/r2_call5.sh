#!/bin/bash
# Round 2 call 5: grouped engine + FFM side stream - ops tests, model tests, bench A/B (CMX_FFM_STREAM)
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py -q > gpurun_out/r2_ops_tests.log 2>&1
echo "ops rc=$?"; tail -5 gpurun_out/r2_ops_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_dp_gpu.py -q -s > gpurun_out/r2_model_tests.log 2>&1
echo "model rc=$?"; grep -E "passed|failed|error|Error|gradient norms outside" gpurun_out/r2_model_tests.log | tail -12
grep "worst gradient cosines" gpurun_out/r2_model_tests.log > gpurun_out/r2_gradient_cosines.txt
for flag in 1 0; do
CMX_FFM_STREAM=$flag timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ffm$flag.json 2> gpurun_out/r2_bench_ffm$flag.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_ffm$flag.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench_ffm$flag.json"))
print("CMX_FFM_STREAM=$flag ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"], "infer", d["inference"]["batch8"])
P
done
