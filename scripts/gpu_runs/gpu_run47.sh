mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/kernels_r1_v3.csv > gpurun_out/bench_v3_n1.json 2> gpurun_out/bench_v3_n1.err
tail -2 gpurun_out/bench_v3_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v3_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','cpu_baseline','clocks')}); print(d['e2e']); print(d['roofline']); print(d['inference'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v3_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_v3_ref.json
