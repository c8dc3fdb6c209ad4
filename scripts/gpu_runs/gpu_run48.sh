mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "im2col or conv" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_48.csv 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NEW', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches_per_step'], d['inference']['batch8']['ms_per_forward'], d['inference']['batch1']['ms_per_forward'])"; done
grep convw gpurun_out/kernels_48.csv
