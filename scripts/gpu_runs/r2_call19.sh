#!/bin/bash
mkdir -p gpurun_out
for thr in 148 296 444 148 296; do
CMX_GEMM_BN64_BELOW=$thr timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench19_$thr.json 2> gpurun_out/r2_bench19_$thr.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench19_$thr.json"))
print("BN64_BELOW=$thr ms/step", d["ms_per_step"], "img/s", d["value"], "infer", d["inference"]["batch8"]["ms_per_forward"], d["inference"]["batch1"]["ms_per_forward"])
P
done
