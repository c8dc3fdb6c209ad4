mkdir -p gpurun_out
timeout 300 python scripts/gemm_microbench.py 2>&1 | tee gpurun_out/gemm_micro.log
ITERS=1 ONLY=0,1 timeout 300 python scripts/gemm_microbench.py > /dev/null 2>&1 && \
ITERS=1 ONLY=0,1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 6 -o gpurun_out/prof_gemm_micro python scripts/gemm_microbench.py > gpurun_out/ncu_micro.log 2>&1
tail -3 gpurun_out/ncu_micro.log
