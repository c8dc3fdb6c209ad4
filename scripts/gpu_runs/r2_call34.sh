mkdir -p gpurun_out
for thr in 148 460 620 148 460; do
CMX_GEMM_BN64_BELOW=$thr timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench34_$thr.json 2> gpurun_out/r2_bench34_$thr.err
python - <<P
import json
d = json.load(open("gpurun_out/r2_bench34_$thr.json"))
print("bn64_below=$thr ms/step", d["ms_per_step"], "img/s", d["value"])
P
done
