#!/bin/bash
# Round 2 call 12: source-level ncu of the stage-1 streaming kernels (LayerNorm fwd/bwd, depthwise conv x3, attention fwd/bwd)
mkdir -p gpurun_out
set -x
K='regex:dwconv_tiled|ln_bwd_v2|ln_fwd_v2|attn_kernel'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 444 -c 6 -o gpurun_out/r2_s1_fwd python scripts/ncu_step.py --steps 1 > gpurun_out/r2_ncu_s1_fwd.log 2>&1
echo "rc=$?"; grep "launches" gpurun_out/r2_ncu_s1_fwd.log | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 658 -c 8 -o gpurun_out/r2_s1_bwd python scripts/ncu_step.py --steps 1 > gpurun_out/r2_ncu_s1_bwd.log 2>&1
echo "rc=$?"
