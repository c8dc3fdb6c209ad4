mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err
tail -3 gpurun_out/bench11.err; python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print('PDL on ', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['e2e'])"
CMX_PDL=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench11b.json 2> gpurun_out/bench11b.err
python -c "
import json; d=json.load(open('gpurun_out/bench11b.json')); print('PDL off', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')})"
