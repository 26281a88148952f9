#!/bin/bash
# round-2 GPU job 1: new 640^2 parity tests + full GPU suite, baseline bench, per-kernel launch lists (train step, NMS)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2j1_pytest.log
tail -5 gpurun_out/r2j1_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/r2j1_per_op.json > gpurun_out/r2j1_bench.log 2>&1; tail -1 gpurun_out/r2j1_bench.log | cut -c1-400
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2j1_train.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j1_train_launches.csv \
  python tools/bench_train.py --bs 8 --steps 1 --warmup 2 --no-graphs > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j1_nms_launches.csv \
  python tools/run_nms.py > gpurun_out/r2j1_nms.log 2>&1
tools/gpu_sanity.sh end
