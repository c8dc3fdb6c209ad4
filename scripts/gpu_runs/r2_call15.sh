#!/bin/bash
mkdir -p gpurun_out
for shp in "19200 320 320" "19200 1280 320" "19200 320 1280" "4800 512 2048"; do echo "=== $shp"; timeout 120 python scripts/gemm_trace.py $shp 2>&1 | head -14; done
