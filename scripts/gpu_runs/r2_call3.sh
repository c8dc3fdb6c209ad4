#!/bin/bash
# Round 2 call 3: full GPU suite on 2 GPUs (DP numerics incl.), gradient-cosine report, per-shape kernel profile
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_gpu_tests.log 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2_gpu_tests.log | tail -5
grep "worst gradient cosines" gpurun_out/r2_gpu_tests.log > gpurun_out/r2_gradient_cosines.txt
for mode in flat ddp; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
      tests/dp_worker.py --mode $mode > gpurun_out/r2_dp_worker_$mode.log 2>&1
  echo "rc=$?"; grep -E "call|DP_WORKER_OK|Error|assert" gpurun_out/r2_dp_worker_$mode.log | tail -8
done
CMX_PROFILE_SHAPES=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r2_shapes.csv > gpurun_out/r2_bench_shapes.json 2> gpurun_out/r2_bench_shapes.err
python - <<P
import torch, time
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")  # 4 GiB
for name, fn, nb in (("fill (write only)", lambda: x.zero_(), 4 * x.numel()), ("sum (read only)", lambda: x.sum(), 4 * x.numel())):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.0f GB/s" % (10 * nb / (e0.elapsed_time(e1) * 1e-3) / 1e9))
P
