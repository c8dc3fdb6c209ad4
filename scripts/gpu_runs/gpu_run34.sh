mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "flat_adamw or graph" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_34.csv > gpurun_out/bench_34.json 2> gpurun_out/bench_34.err; tail -3 gpurun_out/bench_34.err; python -c "
import json; d=json.load(open('gpurun_out/bench_34.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','inference')}); print(d['e2e'])"
timeout 600 python bench.py --no-cpu-baseline --optimizer torch 2>/dev/null | cut -c1-250
