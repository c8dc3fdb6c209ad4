mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v5_n2.json 2> gpurun_out/bench_v5_n2.err
tail -3 gpurun_out/bench_v5_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v5_n2.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'], d['config'].get('grad_allreduce'))"
