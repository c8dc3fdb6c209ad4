# last call of round 2: the whole GPU suite on the final tree (new: pickling / deepcopy after training steps)
mkdir -p gpurun_out
timeout 75 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_gpu_tests_final_check.log
tail -12 gpurun_out/r2_gpu_tests_final_check.log
