timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1t.json > gpurun_out/bench_r1t.log 2>&1; tail -1 gpurun_out/bench_r1t.log | cut -c1-200
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-1500
tools/gpu_sanity.sh end
