mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -12
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -q -p no:cacheprovider 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1c.csv > gpurun_out/bench3.json 2> gpurun_out/bench3.err
tail -3 gpurun_out/bench3.err; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['roofline']); print(d['e2e'])"; head -16 gpurun_out/kernels_r1c.csv
