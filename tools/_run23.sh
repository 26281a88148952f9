timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1o.json > gpurun_out/bench_r1o.log 2>&1; tail -1 gpurun_out/bench_r1o.log | cut -c1-200; tail -1 gpurun_out/bench_r1o.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['roofline']['frac'])"
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "spp" 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-1500
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 74 --launch-count 3 -f -o gpurun_out/ncu_early_layers2 python tools/run_forward.py 2 > gpurun_out/ncu_early2.log 2>&1; tail -1 gpurun_out/ncu_early2.log
tools/gpu_sanity.sh end
