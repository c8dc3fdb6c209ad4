#!/bin/bash
mkdir -p gpurun_out
for flag in 0 1 0 1; do
CMX_ATTN_DKV_RECOMPUTE=$flag timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench16_$flag.json 2> gpurun_out/r2_bench16_$flag.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench16_$flag.json"))
print("RECOMPUTE=$flag ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"])
for k in d["top_kernels"][:12]: print("   ", k)
P
done
