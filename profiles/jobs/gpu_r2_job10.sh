#!/bin/bash
# round-2 GPU job 10 (8 GPUs): DDP check on 8 ranks, bench.py under torchrun at N = 8
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 tests/diag/check_ddp.py > gpurun_out/r2j10_ddp.log 2>&1
grep "graphs=\|DDP_" gpurun_out/r2j10_ddp.log
timeout 900 $TR --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j10_bench_8gpu.log 2> gpurun_out/r2j10_bench_8gpu.err
tail -c 1800 gpurun_out/r2j10_bench_8gpu.log; grep -v "^W\|arn" gpurun_out/r2j10_bench_8gpu.err | tail -3
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader | head -8
