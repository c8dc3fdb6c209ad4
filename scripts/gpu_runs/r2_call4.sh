#!/bin/bash
# Round 2 call 4: grouped (branch-batched) engine - ops tests, model tests, DP tests (2 GPUs), bench + shape profile
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py -q -x > gpurun_out/r2_ops_tests.log 2>&1
echo "ops rc=$?"; tail -5 gpurun_out/r2_ops_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_dp_gpu.py -q -x -s > gpurun_out/r2_model_tests.log 2>&1
echo "model rc=$?"; grep -E "passed|failed|error|Error" gpurun_out/r2_model_tests.log | tail -8
grep "worst gradient cosines" gpurun_out/r2_model_tests.log > gpurun_out/r2_gradient_cosines.txt
CMX_PROFILE_SHAPES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_shapes.csv > gpurun_out/r2_bench_shapes.json 2> gpurun_out/r2_bench_shapes.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_shapes.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench_shapes.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"])
P
