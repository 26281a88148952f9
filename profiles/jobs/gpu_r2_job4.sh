#!/bin/bash
# round-2 GPU job 4: f1/f2/f4 kernels (letterbox, val matching, TTA), nn.Module facade + reference seam, BN kernel retune
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pre_gpu.py tests/test_tta_gpu.py tests/test_val_gpu.py tests/test_zz_reference_seam_gpu.py tests/test_train_gpu.py tests/test_train_layers_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2j4_pytest.log
tail -25 gpurun_out/r2j4_pytest.log
timeout 300 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2j4_train.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j4_train_launches.csv \
  python tools/bench_train.py --bs 8 --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
tools/gpu_sanity.sh end
