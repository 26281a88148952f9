#!/bin/bash
# round-2 GPU job 9 (2 GPUs): DDP hardware check with full logs; phase-decomposed stride-2 dgrad tests + train timing on GPU 0
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29521 tests/diag/check_ddp.py > gpurun_out/r2j9_ddp.log 2>&1
grep -v "Warning\|warn" gpurun_out/r2j9_ddp.log | tail -25
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_layers_gpu.py tests/test_conv_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2j9_pytest.log
tail -6 gpurun_out/r2j9_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/r2j9_train.log
tools/gpu_sanity.sh end
