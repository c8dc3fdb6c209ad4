#!/bin/bash
# attention forward: pipelined x16 TMEM loads, quarter barriers, sum-then-normalised-store passes; parity + standalone timing
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention or attn" > gpurun_out/r2_tests26.log 2>&1
echo "attn tests rc=$?"; tail -5 gpurun_out/r2_tests26.log
timeout 300 python scripts/attn_microbench.py > gpurun_out/r2_attn_microbench_new.txt 2>&1
cat gpurun_out/r2_attn_microbench_new.txt
timeout 900 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_model26.log 2>&1
echo "model rc=$?"; tail -2 gpurun_out/r2_model26.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench26_$i.json 2> gpurun_out/r2_bench26_$i.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench26_$i.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "infer", d["inference"]["batch8"]["ms_per_forward"])
P
done
