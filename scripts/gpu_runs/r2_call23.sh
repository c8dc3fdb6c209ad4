#!/bin/bash
# dwconv 20-wide tiles for narrow maps + fused small-M linear backward: parity, step time, per-shape table
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_ops_gpu.py -q -x > gpurun_out/r2_tests23.log 2>&1
echo "ops tests rc=$?"; tail -3 gpurun_out/r2_tests23.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model23.log 2>&1
echo "model rc=$?"; tail -2 gpurun_out/r2_model23.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench23_$i.json 2> gpurun_out/r2_bench23_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench23_$i.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "launches", d["gpu_launches"])
P
done
timeout 300 python scripts/gemm_shape_replay.py > gpurun_out/r2_shape_replay23.txt 2> gpurun_out/r2_shape_replay23.err
grep "dwconv\|smallm\|^#" gpurun_out/r2_shape_replay23.txt
