timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1s.json > gpurun_out/bench_r1s.log 2>&1; tail -1 gpurun_out/bench_r1s.log | cut -c1-200
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-1500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/bench_train.py --bs 8 --steps 1 --warmup 2 --no-graphs > gpurun_out/train_ncu.log 2>&1; tail -2 gpurun_out/train_ncu.log | cut -c1-300
tools/gpu_sanity.sh end
