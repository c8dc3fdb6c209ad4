mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/tc4.log
tail -8 gpurun_out/tc4.log
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|FAILED|Error|rel diff|assert" | tail -30 | tee gpurun_out/model_tests4.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1b.csv > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -3 gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','roofline')}); print(d['e2e'])"; head -25 gpurun_out/kernels_r1b.csv
