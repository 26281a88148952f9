timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3 | cut -c1-400
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -E "^E  |passed|failed" | head -6 | cut -c1-1000
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/bench_train.py --bs 16 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
tools/gpu_sanity.sh end
