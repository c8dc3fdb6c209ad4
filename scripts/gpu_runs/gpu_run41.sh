mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/kernels_r1_v2.csv > gpurun_out/bench_v2_n1.json 2> gpurun_out/bench_v2_n1.err
tail -2 gpurun_out/bench_v2_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v2_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','cpu_baseline','clocks')}); print(d['e2e']); print(d['roofline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v2_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_v2_ref.json
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain41.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1_v2.csv python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_list41.log 2>&1
tail -1 gpurun_out/plain41.log
gzip -f gpurun_out/launches_r1_v2.csv
timeout 1500 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|attn_kernel|dwconv_tiled|ln_bwd_v2" -s 2160 -c 12 -o gpurun_out/prof_r1_v2_top python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_full41.log 2>&1
tail -1 gpurun_out/ncu_full41.log; ls -la gpurun_out/
