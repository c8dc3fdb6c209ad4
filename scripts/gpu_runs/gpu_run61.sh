# final round-1 artifacts: full GPU test suite, smoke, ncu launch list (+ DRAM bytes) of one eager step, bench (N=1), reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_61.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_61.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python scripts/ncu_step.py --steps 1 > gpurun_out/plain61.log 2>&1 || exit 1
tail -1 gpurun_out/plain61.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r1_final.csv python scripts/ncu_step.py --steps 1 > gpurun_out/ncu_list61.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_list61.log
python scripts/ncu_launch_summary.py gpurun_out/launches_r1_final.csv --from-last convw_pack_multi --note "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, last eager training step (fwd+bwd, no optimizer) of scripts/ncu_step.py, MiT-B2 480x640 batch 8, 1 x B200, final round-1 code" --csv gpurun_out/r1_ncu_launch_list_dram_final.csv --json gpurun_out/ncu_traffic_by_class.json | head -12
cp gpurun_out/ncu_traffic_by_class.json profiles/ncu_traffic_by_class.json
timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/kernels_r1_v5.csv > gpurun_out/bench_v5_n1.json 2> gpurun_out/bench_v5_n1.err
tail -2 gpurun_out/bench_v5_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v5_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','cpu_baseline')}); print(d['e2e']); print(d['roofline']['step'], d['roofline']['traffic'], d['roofline']['frac']); print({k:(v.get('img_s'),v.get('ms_per_forward',v.get('ms_per_image'))) for k,v in d['inference'].items()})"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v5_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench_v5_ref.json
