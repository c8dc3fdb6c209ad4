mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k sliding 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_40.json 2> gpurun_out/bench_40.err; tail -3 gpurun_out/bench_40.err; python -c "
import json; d=json.load(open('gpurun_out/bench_40.json')); print({k:d[k] for k in ('value','ms_per_step')}); print(json.dumps(d['inference'],indent=1)[:1800])"
