mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "dwconv" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_30.csv > gpurun_out/bench_30.json 2> gpurun_out/bench_30.err; python -c "
import json; d=json.load(open('gpurun_out/bench_30.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','inference')}); print(d['e2e'])"
head -14 gpurun_out/kernels_30.csv
