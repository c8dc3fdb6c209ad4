mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_25.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_25.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default_25.json 2> gpurun_out/bench_default_25.err; echo "bench rc=$?"; wc -l gpurun_out/bench_default_25.json; cut -c1-900 gpurun_out/bench_default_25.json
