mkdir -p gpurun_out
timeout 900 python tests/tools/gpu_debug_model.py b2_small b0_odd b4_small 2>&1 | grep -v "rel-L2 [0-9.e-]*$" > gpurun_out/debug_model.log
tail -120 gpurun_out/debug_model.log
