mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -p no:cacheprovider 2>&1 | tail -12
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -6
CMX_PROFILE_SHAPES=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_shapes2.csv > gpurun_out/bench5.json 2> gpurun_out/bench5.err
tail -3 gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print(d['e2e'])"; grep -E "gemm_tc_(fwd|dgrad)_(153600|19200|38400)" gpurun_out/kernels_shapes2.csv | head -20
