#!/bin/bash
# Round 2 call 11: source-level ncu of the three hot epilogue bodies + the split-K weight-gradient path
mkdir -p gpurun_out
set -x
ONLY=9,10,14,w1 ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 12 -o gpurun_out/r2_gemm_epi2 python scripts/gemm_microbench.py > gpurun_out/r2_ncu_gemm2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_gemm2.log
