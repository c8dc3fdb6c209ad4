mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/kernels_r1_final.csv > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
tail -3 gpurun_out/bench_final_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','inference','cpu_baseline','clocks')}); print(d['e2e']); print(d['roofline'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json | cut -c1-600
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1b.csv python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_list2.log 2>&1
tail -2 gpurun_out/plain2.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_kernel" -s 1700 -c 16 -o gpurun_out/prof_r1b_gemm_attn python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log; ls -la gpurun_out/*.ncu-rep
