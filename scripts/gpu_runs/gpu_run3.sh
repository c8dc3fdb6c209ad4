mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/model_tests.log
tail -15 gpurun_out/model_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/kernels_r1.csv > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json; cat gpurun_out/kernels_r1.csv
