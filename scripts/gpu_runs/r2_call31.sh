#!/bin/bash
# evaluation driver: streaming accumulators (no per-image read-back) - parity + bench line
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_eval_gpu.py -q -x > gpurun_out/r2_tests31.log 2>&1
echo "eval tests rc=$?"; tail -3 gpurun_out/r2_tests31.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench31.json 2> gpurun_out/r2_bench31.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench31.json"))
print("ms/step", d["ms_per_step"])
for k, v in d["inference"].items():
    if k.startswith("sliding"): print(k, round(v["ms_per_image"], 2), "ms/img")
P
