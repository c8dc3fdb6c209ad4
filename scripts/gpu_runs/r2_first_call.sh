#!/bin/bash
# Round-2 first GPU call: model-level validation and A/B of the flash-style attention backward (csrc/attention_dkv.cu;
# its kernel-level parity test already passed on B200 in round 1).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_runs/r2_first_call.sh'
# Every step has its own timeout (the mbarrier watchdog turns protocol bugs into traps, not hangs).
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -q -k dkv_recompute > gpurun_out/r2_dkv_ops.log 2>&1
rc=$?
tail -15 gpurun_out/r2_dkv_ops.log
if [ $rc -ne 0 ]; then echo "experimental kernels failed parity (rc=$rc): not benchmarking the flagged path"; exit 0; fi
CMX_ATTN_DKV_RECOMPUTE=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -x > gpurun_out/r2_dkv_model.log 2>&1
tail -5 gpurun_out/r2_dkv_model.log
for flag in 0 1 0 1; do
  CMX_ATTN_DKV_RECOMPUTE=$flag timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r2_bench_dkv$flag.json 2> gpurun_out/r2_bench_dkv$flag.err
  python - <<P
import json
d = json.load(open("gpurun_out/r2_bench_dkv$flag.json"))
print("CMX_ATTN_DKV_RECOMPUTE=$flag", d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches_per_step"])
P
done
