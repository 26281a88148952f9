timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-1200
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-700
Y3_WGRAD_TC=0 timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-500
timeout 200 python tools/bench_train.py --bs 16 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-500
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1u.log 2>&1; tail -1 gpurun_out/bench_r1u.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks'])"
tools/gpu_sanity.sh end
