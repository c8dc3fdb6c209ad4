mkdir -p gpurun_out
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain52.log 2>&1 || exit 1
tail -1 gpurun_out/plain52.log
timeout 1500 ncu --set full --clock-control none -k regex:"attn_kernel|dwconv_tiled|ln_bwd_v2" -s 270 -c 9 -o gpurun_out/prof_r1_v3_membound python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_full52.log 2>&1
tail -1 gpurun_out/ncu_full52.log; ls -la gpurun_out/*.ncu-rep
