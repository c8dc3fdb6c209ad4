#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention or attn" > gpurun_out/r2_tests21.log 2>&1
echo "attn tests rc=$?"; tail -3 gpurun_out/r2_tests21.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model21.log 2>&1
echo "model rc=$?"; tail -2 gpurun_out/r2_model21.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench21_$i.json 2> gpurun_out/r2_bench21_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench21_$i.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "infer", d["inference"]["batch8"]["ms_per_forward"])
for k in d["roofline"]["kernels"]:
    if "attn" in k["kernel"]: print("   ", k["kernel"], k["avg_us"], k["frac"])
P
done
timeout 600 python bench.py --config b4_pst900 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench21_b4.json 2> gpurun_out/r2_bench21_b4.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench21_b4.json"))
print("b4 ms/step", d["ms_per_step"], "img/s", d["value"])
for k in d["top_kernels"]: print("   ", k)
P
