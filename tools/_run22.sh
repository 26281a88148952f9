timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_nms_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1n.json > gpurun_out/bench_r1n.log 2>&1; tail -1 gpurun_out/bench_r1n.log | cut -c1-200; tail -1 gpurun_out/bench_r1n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['roofline']['frac'], d.get('nms',{}).get('conf0.25_iou0.45_single'))"
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -5 | cut -c1-600
tools/gpu_sanity.sh end
