mkdir -p gpurun_out
E=rgbx_semantic_segmentation_b200/engine.py
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NEW', d['ms_per_step'], d['e2e']['ms_per_step'], d['inference']['batch8']['ms_per_forward'])"
cp $E /tmp/engine_new.py; cp scripts/_engine_old.py $E
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('OLD', d['ms_per_step'], d['e2e']['ms_per_step'], d['inference']['batch8']['ms_per_forward'])"
cp /tmp/engine_new.py $E
done
