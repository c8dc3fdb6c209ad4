#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests14.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2_tests14.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench14_$i.json 2> gpurun_out/r2_bench14_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench14_$i.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "infer", d["inference"]["batch8"])
for k in d["top_kernels"][:10]: print("   ", k)
P
done
