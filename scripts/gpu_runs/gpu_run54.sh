mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_54.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_54.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/kernels_r1_v4.csv > gpurun_out/bench_v4_n1.json 2> gpurun_out/bench_v4_n1.err
tail -2 gpurun_out/bench_v4_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v4_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','cpu_baseline')}); print(d['e2e']); print(d['roofline']['step']); print({k:(v.get('img_s'),v.get('ms_per_forward',v.get('ms_per_image'))) for k,v in d['inference'].items()})"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v4_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_v4_ref.json
