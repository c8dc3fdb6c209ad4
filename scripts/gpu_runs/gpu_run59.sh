mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "upsample" 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --profile-out gpurun_out/kernels_59.csv 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NEW', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
grep "upsample\|colstats\|bn_" gpurun_out/kernels_59.csv
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain59.log 2>&1 || exit 1
tail -1 gpurun_out/plain59.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"upsample_sum|upsample_bwd_multi|bn_bwd_reduce_v8|bn_bwd_apply_v8|colstats|bn_apply_kernel" -s 92 -c 6 -o gpurun_out/prof_r1_decoder python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_full59.log 2>&1
tail -1 gpurun_out/ncu_full59.log; ls -la gpurun_out/*.ncu-rep
