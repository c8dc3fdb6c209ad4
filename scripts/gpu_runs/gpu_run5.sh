mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -q -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider -k "learns" 2>&1 | tail -5
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1.csv python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_list.log 2>&1
cat gpurun_out/plain.log | tail -4
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1990 -c 14 -o gpurun_out/prof_gemm_tc_r1 python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv_bwd_pre|ln_bwd_kernel|colsum_kernel" -s 772 -c 6 -o gpurun_out/prof_membound_r1 python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/ | tail; wc -l gpurun_out/launches_r1.csv; tail -3 gpurun_out/ncu_full1.log gpurun_out/ncu_full2.log
