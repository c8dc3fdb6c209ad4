#!/bin/bash
# Round 2 call 13: trimmed attention softmax / dwconv addressing / single-MUFU exp2: tests + bench + b4 bench + shape profile
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_tc_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_tests13.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_tests13.log
CMX_PROFILE_SHAPES=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_shapes13.csv > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench13b.json 2> gpurun_out/r2_bench13b.err
timeout 600 python bench.py --config b4_pst900 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench13_b4.json 2> gpurun_out/r2_bench13_b4.err
python - <<P
import json
for f in ("gpurun_out/r2_bench13.json", "gpurun_out/r2_bench13b.json", "gpurun_out/r2_bench13_b4.json"):
    d = json.load(open(f))
    print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "infer", d["inference"])
    for k in d["top_kernels"][:8]: print("   ", k)
P
