mkdir -p gpurun_out
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v5_n$N.json 2> gpurun_out/bench_v5_n$N.err
tail -2 gpurun_out/bench_v5_n$N.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_v5_n$N.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])"
