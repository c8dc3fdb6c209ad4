mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
tail -3 gpurun_out/bench12.err; python -c "
import json; d=json.load(open('gpurun_out/bench12.json')); print('dual-stream on ', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['e2e'])"
CMX_DUAL_STREAM=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench12b.json 2> gpurun_out/bench12b.err
python -c "
import json; d=json.load(open('gpurun_out/bench12b.json')); print('dual-stream off', {k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')})"
