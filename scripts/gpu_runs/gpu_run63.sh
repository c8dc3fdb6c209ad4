mkdir -p gpurun_out
for i in 1 2; do CMX_HP_STREAMS=1 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('HP_STREAMS', d['ms_per_step'], d['e2e']['ms_per_step'], d['last_loss'], d['inference']['batch8']['ms_per_forward'], d['inference']['batch1']['ms_per_forward'])"; done
for i in 1; do timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('DEFAULT', d['ms_per_step'], d['e2e']['ms_per_step'], d['last_loss'], d['inference']['batch8']['ms_per_forward'], d['inference']['batch1']['ms_per_forward'])"; done
timeout 600 python scripts/bench_b4_pst900.py 2>&1 | tail -6
