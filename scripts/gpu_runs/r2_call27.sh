#!/bin/bash
# full GPU suite on the current code + bench + per-shape table
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/r2_tests27.log 2>&1
echo "gpu suite rc=$?"; tail -4 gpurun_out/r2_tests27.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench27_$i.json 2> gpurun_out/r2_bench27_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench27_$i.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "infer", d["inference"]["batch8"]["ms_per_forward"], "launches", d["gpu_launches"])
P
done
timeout 300 python scripts/gemm_shape_replay.py > gpurun_out/r2_shape_replay27.txt 2> gpurun_out/r2_shape_replay27.err
grep "frm_rectify\|im2col_nchw\|smallm\|^#" gpurun_out/r2_shape_replay27.txt
