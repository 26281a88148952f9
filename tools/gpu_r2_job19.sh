#!/bin/bash
# one short call: on-device A/B of the decode / BatchNorm kernel variants (bit-equality + timing), see tests/diag/ab_shot.py
set -u
mkdir -p gpurun_out
timeout 175 python tests/diag/ab_shot.py --budget 115 > gpurun_out/ab_shot.log 2>&1
tail -4 gpurun_out/ab_shot.log
