mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_31.csv > gpurun_out/bench_31.json 2> gpurun_out/bench_31.err; python -c "
import json; d=json.load(open('gpurun_out/bench_31.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','inference')}); print(d['e2e'])"
CMX_GEMM_TILE_POLICY=pad timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-250
head -5 gpurun_out/kernels_31.csv
