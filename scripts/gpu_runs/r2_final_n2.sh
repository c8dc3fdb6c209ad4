#!/bin/bash
# Round 2 final: 2 GPUs of one box, one rank per GPU over NCCL (as the driver launches it)
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_runs/r2_final_n2.sh'
mkdir -p gpurun_out
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
echo "bench rc=$?"
timeout 900 python -m pytest tests/test_dp_gpu.py -q > gpurun_out/r2_dp_tests_2gpu.log 2>&1
echo "dp tests rc=$?"; tail -3 gpurun_out/r2_dp_tests_2gpu.log
for mode in reference flat ddp; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/dp_worker.py --mode $mode 2>&1 | grep -E "call [0-9]|reference DDP|DP_WORKER_OK" > gpurun_out/r2_dp_worker_$mode.txt
  cat gpurun_out/r2_dp_worker_$mode.txt
done
python - <<P
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*n[248]*.json")):
    try:
        d = json.load(open(f))
        print(f, "n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "img/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "infer", {k: round(v["img_s"], 1) for k, v in d["inference"].items() if k.startswith("batch")}, "ddp", d.get("ddp_stock") and round(d["ddp_stock"]["img_s"], 1), "norm", d["config"].get("norm_layer"), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
P
