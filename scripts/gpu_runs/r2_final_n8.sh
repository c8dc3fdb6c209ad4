#!/bin/bash
# Round 2 final: 8 GPUs of one box, one rank per GPU over NCCL (as the driver launches it)
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_runs/r2_final_n8.sh'
mkdir -p gpurun_out
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
echo "bench rc=$?"
if [ "$1" = "b4" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --config b4_pst900 --steps 10 --warmup 3 --no-ddp-compare > gpurun_out/r2_bench_b4_pst900_n8.json 2> gpurun_out/r2_bench_b4_pst900_n8.err
echo "b4 n8 rc=$?"
fi
python - <<P
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*n8*.json")):
    try:
        d = json.load(open(f))
        print(f, "n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "img/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "infer", {k: round(v["img_s"], 1) for k, v in d["inference"].items() if k.startswith("batch")}, "ddp", d.get("ddp_stock") and round(d["ddp_stock"]["img_s"], 1), "norm", d["config"].get("norm_layer"), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
P
