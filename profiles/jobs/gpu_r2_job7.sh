#!/bin/bash
# round-2 GPU job 7: yolov3-tiny training (pools backward, 16-channel layers), NMS final composition, seam test
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_gpu.py tests/test_train_layers_gpu.py tests/test_zz_reference_seam_gpu.py tests/test_nms_gpu.py tests/test_pipeline_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider --tb=long 2>&1 | tail -120 > gpurun_out/r2j7_pytest.log
tail -8 gpurun_out/r2j7_pytest.log
for c in "0.25 0.45 0" "0.001 0.6 0" "0.25 0.45 1" "0.001 0.6 1"; do set -- $c; timeout 120 python tools/run_nms.py --conf $1 --iou $2 --ml $3 --iters 10 2>&1 | tail -1; done | tee gpurun_out/r2j7_nms.log
timeout 300 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 --cfg yolov3-tiny.yaml 2>&1 | tail -1 | cut -c1-700 | tee gpurun_out/r2j7_train_tiny.log
tools/gpu_sanity.sh end
