#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench20.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench20.json"))
print("ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"])
r = d["roofline"]
print({k: r[k] for k in ("kernel", "bound", "achieved", "achieved_event_pairs", "frac", "avg_us", "avg_us_event_pairs", "event_floor_us", "traffic")})
for k in r["kernels"]: print("  ", k)
print(r["step"])
P
