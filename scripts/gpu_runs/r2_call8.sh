#!/bin/bash
# Round 2 call 8: where does the GEMM epilogue time go?  full clock64 trace + ncu source-level capture of the write-heavy shape
mkdir -p gpurun_out
set -x
for shp in "307200 256 64" "307200 64 256"; do timeout 120 python scripts/gemm_trace.py $shp > gpurun_out/r2_gemm_trace_$(echo $shp | tr ' ' 'x').txt 2>&1; done
ONLY=9,10,14 ITERS=5 timeout 120 python scripts/gemm_microbench.py > gpurun_out/r2_gemm_micro.txt 2>&1
ONLY=9 ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r2_gemm_epi python scripts/gemm_microbench.py > gpurun_out/r2_ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_gemm.log
cat gpurun_out/r2_gemm_micro.txt
