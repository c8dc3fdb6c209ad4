#!/bin/bash
# Round 2 call 7: key-chunked attention (ops + b4 model tests), uint8 input pipeline, b4_pst900 bench, GEMM trace
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "chunked or dkv_recompute or attention" > gpurun_out/r2_attn_tests.log 2>&1
echo "attn rc=$?"; tail -4 gpurun_out/r2_attn_tests.log
timeout 600 python -m pytest tests/test_input_pipeline_gpu.py tests/test_eval_gpu.py -q > gpurun_out/r2_pipe_tests.log 2>&1
echo "pipe rc=$?"; tail -6 gpurun_out/r2_pipe_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -s -k "b4 or golden" > gpurun_out/r2_b4_tests.log 2>&1
echo "b4 rc=$?"; grep -E "passed|failed|error|Error|gradient norms outside|FAILED" gpurun_out/r2_b4_tests.log | tail -8
timeout 600 python bench.py --config b4_pst900 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b4.json 2> gpurun_out/r2_bench_b4.err
echo "bench b4 rc=$?"; tail -3 gpurun_out/r2_bench_b4.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench_b4.json"))
print("b4_pst900 ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"], "infer", d["inference"]["batch%d" % d["config"]["per_gpu_batch"]])
for k in d["top_kernels"]: print("   ", k)
P
for shp in "307200 256 64" "307200 64 256" "307200 64 64"; do timeout 120 python scripts/gemm_trace.py $shp 2>&1 | tail -12; done
