mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.log 2>&1
timeout 1200 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider 2>&1 | tail -120 > gpurun_out/ops.log
for t in test_tc_forward_plain test_tc_forward_epilogue test_tc_dgrad_b_mn_major test_tc_wgrad_mn_major_splitk test_auto_dispatch_prefers_tc; do
  echo "=== $t" >> gpurun_out/tc.log
  timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q -p no:cacheprovider -k $t 2>&1 | tail -60 >> gpurun_out/tc.log
done
tail -5 gpurun_out/ops.log; grep -E "passed|failed|error" gpurun_out/tc.log
