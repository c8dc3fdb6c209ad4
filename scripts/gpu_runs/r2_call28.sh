#!/bin/bash
# ncu of the TMA-staged depthwise conv kernels (stage 1 and stage 3)
mkdir -p gpurun_out
set -x
ITERS=1 ONLY=0,2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tma -c 6 -o /tmp/dw2 python scripts/dwconv_microbench.py > gpurun_out/r2_ncu_dw2.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/dw2.ncu-rep --page raw --csv > gpurun_out/r2_dw2_raw.csv 2>/dev/null
python scripts/ncu_brief.py /tmp/dw2.ncu-rep
