#!/bin/bash
# Round 2 call 6 (2 GPUs): full GPU suite (eval-driver kernels, b4 golden, DP numerics incl. the 2-process NCCL tests), bench N=1 / N=2
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|error|Error|gradient norms outside|FAILED" gpurun_out/r2_gpu_tests.log | tail -15
grep "worst gradient cosines" gpurun_out/r2_gpu_tests.log > gpurun_out/r2_gradient_cosines.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench2 rc=$?"; grep -v "^W\|NCCL\|^\[W" gpurun_out/r2_bench_n2.err | tail -5
python - <<P
import json
for f in ("gpurun_out/r2_bench_n1.json", "gpurun_out/r2_bench_n2.json"):
    try:
        d = json.load(open(f))
        print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"], "infer", d["inference"]["batch8"], "ddp", d.get("ddp_stock"), "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
P
