#!/bin/bash
# one short call: on-device A/B of the decode / BatchNorm kernel variants (bit-equality + timing), see tests/diag/ab_shot.py;
# then, if the call still has time, the per-layer training tests and the model tests with both variants switched on
set -u
mkdir -p gpurun_out
timeout 150 python tests/diag/ab_shot.py --budget 100 > gpurun_out/ab_shot.log 2>&1
tail -3 gpurun_out/ab_shot.log
Y3_BN_ASYNC=1 Y3_DECODE2=1 timeout 120 python -m pytest tests/test_train_layers_gpu.py tests/test_bench_config_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/ab_pytest.log 2>&1
tail -3 gpurun_out/ab_pytest.log
