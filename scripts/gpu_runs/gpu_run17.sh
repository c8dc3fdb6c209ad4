mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -p no:cacheprovider -k "fused_attention" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1f.csv > gpurun_out/bench8.json 2> gpurun_out/bench8.err
tail -3 gpurun_out/bench8.err; python -c "
import json; d=json.load(open('gpurun_out/bench8.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['e2e'])"; head -14 gpurun_out/kernels_r1f.csv; grep attn gpurun_out/kernels_r1f.csv
