mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "upsample or ce_ or im2col or focal or conv" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_57.csv 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NEW', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
grep "col2im\|im2col\|upsample\|ce_" gpurun_out/kernels_57.csv
