mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q 2>&1 | tail -3
timeout 250 python scripts/gemm_shapes_bench.py > gpurun_out/gemm_shapes_auto2.log 2>&1; head -16 gpurun_out/gemm_shapes_auto2.log
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_32.csv > gpurun_out/bench_32.json 2> gpurun_out/bench_32.err; python -c "
import json; d=json.load(open('gpurun_out/bench_32.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','inference')}); print(d['e2e'])"
head -5 gpurun_out/kernels_32.csv
