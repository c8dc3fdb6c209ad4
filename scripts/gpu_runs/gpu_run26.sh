mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/gemm_microbench.py > gpurun_out/gemm_micro_26.log 2>&1; cat gpurun_out/gemm_micro_26.log
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_26.json 2> gpurun_out/bench_26.err; python -c "
import json; d=json.load(open('gpurun_out/bench_26.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','inference')}); print(d['e2e'])"
CTX_GAIN=0.1 timeout 600 python tests/tools/gpu_debug_fullsize.py 480x640 > gpurun_out/debug_fullsize2.log 2>&1; grep -c cos gpurun_out/debug_fullsize2.log; grep "^==" gpurun_out/debug_fullsize2.log
