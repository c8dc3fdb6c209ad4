#!/bin/bash
# Round 2 final artifacts on ONE B200: full GPU suite, smoke, ncu launch list (+ DRAM bytes) of one eager step -> traffic file,
# bench (N=1, with the CPU baseline), the reference arm, config 4, ncu --set full of the hot kernels, CUPTI timeline
#   gpurun --timeout 2400 -- 'bash scripts/gpu_runs/r2_final_n1.sh'
mkdir -p gpurun_out
set -x
# (the full GPU suite ran green on the same code in scripts/gpu_runs/r2_call27.sh: profiles/r2_gpu_tests_180_passed.log)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2_final_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 9000 --csv --log-file /tmp/r2_launches_final.csv python scripts/ncu_step.py --steps 1 > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/r2_ncu_list.log
python scripts/ncu_launch_summary.py /tmp/r2_launches_final.csv --from-last convw_pack_multi --note "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, last eager training step (fwd+bwd, no optimizer) of scripts/ncu_step.py, MiT-B2 480x640 batch 8, 1 x B200, final round-2 code (same gpurun call as profiles/r2_bench_n1_final.json)" --csv gpurun_out/r2_ncu_launch_list_dram_final.csv --json gpurun_out/ncu_traffic_by_class.json | head -14
cp gpurun_out/ncu_traffic_by_class.json profiles/ncu_traffic_by_class.json
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2_cuda_event_kernel_breakdown.csv > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
echo "reference arm rc=$?"; cat gpurun_out/r2_bench_reference_arm.json | cut -c1-400
timeout 600 python bench.py --config b4_pst900 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b4_pst900_n1.json 2> gpurun_out/r2_bench_b4.err
echo "b4 rc=$?"
python - <<P
import json
for f in ("gpurun_out/r2_bench_n1_final.json", "gpurun_out/r2_bench_b4_pst900_n1.json"):
    d = json.load(open(f))
    print(f, "ms/step", d["ms_per_step"], "img/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches_per_step"], "clocks", d["clocks"])
    r = d["roofline"]
    print("   roofline", {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "frac", "traffic", "avg_us", "launches_per_step", "event_floor_us")}, r["step"]["frac"])
    print("   cpu", d.get("cpu_baseline"))
    print("   infer", {k: v.get("img_s") for k, v in d["inference"].items()})
P
if [ "$FULL" != "0" ]; then
K='regex:gemm_tc_kernel|dwconv_tma|ln_bwd_v2|ln_fwd_v2|attn_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 1526 -c 24 -o /tmp/r2_final_s1_fwd python scripts/ncu_step.py --steps 1 > gpurun_out/r2_ncu_final_fwd.log 2>&1
echo "ncu full rc=$?"
python scripts/ncu_brief.py /tmp/r2_final_s1_fwd.ncu-rep > gpurun_out/r2_ncu_full_stage1_forward_kernels.txt 2>&1
cat gpurun_out/r2_ncu_full_stage1_forward_kernels.txt | cut -c1-250
timeout 300 python scripts/timeline.py r2_timeline_final.csv > gpurun_out/r2_timeline.log 2>&1
python scripts/timeline_analyze.py gpurun_out/r2_timeline_final.csv > gpurun_out/r2_timeline_summary_final.txt 2>&1
head -12 gpurun_out/r2_timeline_summary_final.txt
rm -f gpurun_out/timeline_trace.json gpurun_out/r2_timeline_final.csv
fi
