#!/bin/bash
# Round 2: data-parallel numerics on 2 GPUs (SyncBatchNorm, FlatDataParallel, stock DDP, reference DDP+SyncBN vs oracle)
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_runs/r2_dp.sh'
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_dp_tests.log 2>&1
tail -15 gpurun_out/r2_dp_tests.log
for mode in reference flat ddp; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
      tests/dp_worker.py --mode $mode > gpurun_out/r2_dp_worker_$mode.log 2>&1
  echo "rc=$?"; grep -v "^W\|^\[W\|NCCL" gpurun_out/r2_dp_worker_$mode.log | tail -12
done
