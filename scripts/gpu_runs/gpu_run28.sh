mkdir -p gpurun_out
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain28.log 2>&1 || exit 1
tail -1 gpurun_out/plain28.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_v2" -s 92 -c 8 -o gpurun_out/prof_r1c_lnbwd python scripts/ncu_step.py --steps 1 > gpurun_out/ncu28a.log 2>&1
tail -1 gpurun_out/ncu28a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv_tiled" -s 96 -c 12 -o gpurun_out/prof_r1c_dwconv python scripts/ncu_step.py --steps 1 > gpurun_out/ncu28b.log 2>&1
tail -1 gpurun_out/ncu28b.log
ls -la gpurun_out/*.ncu-rep
