mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "frm" 2>&1 | tail -5 > gpurun_out/ops2.log
timeout 900 python scripts/gpu_debug_model.py b2_small b0_odd > gpurun_out/debug_model.log 2>&1
tail -5 gpurun_out/ops2.log; tail -100 gpurun_out/debug_model.log
