#!/bin/bash
# round-2 GPU job 8 (2 GPUs): overlapped DDP exchange, SyncBatchNorm, bench.py under torchrun
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/diag/check_ddp.py 2>&1 | grep -v "^W\|Warning\|warn" | tail -8 | tee gpurun_out/r2j8_ddp.log
timeout 400 $TR --master-port 29512 tests/diag/check_syncbn.py 2>&1 | grep -v "^W\|Warning\|warn" | tail -4 | tee gpurun_out/r2j8_syncbn.log
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j8_bench_2gpu.log 2> gpurun_out/r2j8_bench_2gpu.err
tail -c 2500 gpurun_out/r2j8_bench_2gpu.log; tail -5 gpurun_out/r2j8_bench_2gpu.err
tools/gpu_sanity.sh end
