timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -E "^E  |passed|failed" | head -8 | cut -c1-1200
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/bench_train.py --bs 16 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches3.csv python tools/bench_train.py --bs 8 --steps 1 --warmup 2 --no-graphs > gpurun_out/train_ncu3.log 2>&1
tools/gpu_sanity.sh end
