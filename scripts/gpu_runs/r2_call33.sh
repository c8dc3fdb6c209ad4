#!/bin/bash
# small-M linear kernels with deeper load batches: parity, per-shape lines, step time
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_tests33.log 2>&1
echo "ops+model tests rc=$?"; tail -3 gpurun_out/r2_tests33.log
timeout 300 python scripts/gemm_shape_replay.py > gpurun_out/r2_shape_replay33.txt 2> gpurun_out/r2_shape_replay33.err
grep "smallm\|^#" gpurun_out/r2_shape_replay33.txt
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench33_$i.json 2> gpurun_out/r2_bench33_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench33_$i.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"])
P
done
