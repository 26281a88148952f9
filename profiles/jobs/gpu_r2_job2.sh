#!/bin/bash
# round-2 GPU job 2: refactored training path (flat store, two-stage reductions, fused optimizer), new bench legs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_layers_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2j2_pytest_train.log
tail -12 gpurun_out/r2j2_pytest_train.log
timeout 300 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/r2j2_train.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j2_train_launches.csv \
  python tools/bench_train.py --bs 8 --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j2_bench.log 2> gpurun_out/r2j2_bench.err; tail -c 3000 gpurun_out/r2j2_bench.log; tail -5 gpurun_out/r2j2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j2_bench_ref.log 2>&1; tail -c 1500 gpurun_out/r2j2_bench_ref.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_train_gpu.py --deselect tests/test_train_layers_gpu.py 2>&1 | tail -8 > gpurun_out/r2j2_pytest_rest.log
tail -5 gpurun_out/r2j2_pytest_rest.log
tools/gpu_sanity.sh end
