mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -k "mfnet or learns" 2>&1 | grep -E "^E  |passed|failed" | head -20
CMX_PROFILE_SHAPES=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_shapes.csv > gpurun_out/bench4.json 2> gpurun_out/bench4.err
tail -3 gpurun_out/bench4.err; head -60 gpurun_out/kernels_shapes.csv
