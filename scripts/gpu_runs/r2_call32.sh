#!/bin/bash
# fast-division im2col / col2im: parity, per-shape table; ncu of the small-M linear kernels
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_tests32.log 2>&1
echo "ops+model tests rc=$?"; tail -3 gpurun_out/r2_tests32.log
timeout 300 python scripts/gemm_shape_replay.py > gpurun_out/r2_shape_replay32.txt 2> gpurun_out/r2_shape_replay32.err
grep "im2col\|col2im\|smallm\|^#" gpurun_out/r2_shape_replay32.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench32.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"])
P
timeout 300 ncu --set full --clock-control none -k regex:smallm -s 18 -c 18 -o /tmp/smallm python scripts/ncu_step.py --steps 1 > gpurun_out/r2_ncu_smallm.log 2>&1
python scripts/ncu_brief.py /tmp/smallm.ncu-rep | cut -c1-260
