#!/bin/bash
# 16-warp dq / dkv kernels: ops parity, model tests under the recompute backward, A/B of the step time
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention or attn" > gpurun_out/r2_tests22.log 2>&1
echo "attn tests rc=$?"; tail -3 gpurun_out/r2_tests22.log
CMX_ATTN_DKV_RECOMPUTE=1 timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model22.log 2>&1
echo "model(recompute) rc=$?"; tail -2 gpurun_out/r2_model22.log
for mode in 0 1 0 1; do
CMX_ATTN_DKV_RECOMPUTE=$mode timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench22_$mode.json 2> gpurun_out/r2_bench22_$mode.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench22_$mode.json"))
print("recompute=$mode ms/step", d["ms_per_step"], "img/s", d["value"], "launches", d["gpu_launches"])
for k in d["roofline"]["kernels"]:
    if "attn" in k["kernel"]: print("   ", k["kernel"], k["avg_us"], k["frac"])
P
done
for mode in 0 1; do
CMX_ATTN_DKV_RECOMPUTE=$mode timeout 600 python bench.py --config b4_pst900 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench22_b4_$mode.json 2> gpurun_out/r2_bench22_b4_$mode.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench22_b4_$mode.json"))
print("b4 recompute=$mode ms/step", d["ms_per_step"], "img/s", d["value"])
P
done
