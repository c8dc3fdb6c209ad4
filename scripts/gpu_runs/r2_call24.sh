#!/bin/bash
# depthwise conv: grid = whole waves, 2-op bf16 unpack; standalone A/B + ncu (pipes, stalls, source) of the stage-1 kernels
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "dwconv" > gpurun_out/r2_tests24.log 2>&1
echo "dwconv tests rc=$?"; tail -2 gpurun_out/r2_tests24.log
for wv in 0 1 2 3; do
echo "== CMX_DWCONV_WAVES=$wv"
CMX_DWCONV_WAVES=$wv timeout 300 python scripts/dwconv_microbench.py
done > gpurun_out/r2_dwconv_microbench.txt 2>&1
cat gpurun_out/r2_dwconv_microbench.txt
ITERS=1 ONLY=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tiled -c 3 -o /tmp/dw python scripts/dwconv_microbench.py > gpurun_out/r2_ncu_dw.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/dw.ncu-rep --page raw --csv > gpurun_out/r2_dw_raw.csv 2>/dev/null
ncu -i /tmp/dw.ncu-rep --page source --csv > gpurun_out/r2_dw_source.csv 2>/dev/null
ls -la gpurun_out/r2_dw_*.csv
python scripts/ncu_brief.py /tmp/dw.ncu-rep
