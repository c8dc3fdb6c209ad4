mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/gemm_microbench.py 2>&1 | tail -9
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
CMX_PROFILE_SHAPES=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1d.csv > gpurun_out/bench6.json 2> gpurun_out/bench6.err
tail -3 gpurun_out/bench6.err; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['e2e'])"; head -14 gpurun_out/kernels_r1d.csv
