mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider 2>&1 | grep -E "^E  .*(Error|mismatch)|passed|failed" | head -20
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1h.csv > gpurun_out/bench10.json 2> gpurun_out/bench10.err
tail -3 gpurun_out/bench10.err; python -c "
import json; d=json.load(open('gpurun_out/bench10.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['e2e'])"; head -24 gpurun_out/kernels_r1h.csv
