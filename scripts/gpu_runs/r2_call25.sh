#!/bin/bash
# TMA-staged double-buffered depthwise conv: parity, standalone A/B against the cp.async kernel, step time
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "dwconv" > gpurun_out/r2_tests25.log 2>&1
echo "dwconv tests rc=$?"; tail -5 gpurun_out/r2_tests25.log
for tma in 0 1; do
echo "== CMX_DWCONV_TMA=$tma"
CMX_DWCONV_TMA=$tma timeout 300 python scripts/dwconv_microbench.py
done > gpurun_out/r2_dwconv_microbench_tma.txt 2>&1
cat gpurun_out/r2_dwconv_microbench_tma.txt
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model25.log 2>&1
echo "model rc=$?"; tail -2 gpurun_out/r2_model25.log
for tma in 0 1 1; do
CMX_DWCONV_TMA=$tma timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench25_$tma.json 2> gpurun_out/r2_bench25_$tma.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench25_$tma.json"))
print("tma=$tma ms/step", d["ms_per_step"], "img/s", d["value"], "launches", d["gpu_launches"])
P
done
