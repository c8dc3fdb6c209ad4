#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 300 python scripts/timeline.py r2_timeline.csv > gpurun_out/r2_timeline.log 2>&1
tail -3 gpurun_out/r2_timeline.log
python scripts/timeline_analyze.py gpurun_out/r2_timeline.csv > gpurun_out/r2_timeline_summary.txt 2>&1
cat gpurun_out/r2_timeline_summary.txt
rm -f gpurun_out/timeline_trace.json
