#!/bin/bash
# round-2 GPU job 5: re-run the failing seam / TTA tests with full tracebacks, NMS v2 tuning, finalize unrolling
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tta_gpu.py tests/test_zz_reference_seam_gpu.py tests/test_nms_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider --tb=long 2>&1 | tail -150 > gpurun_out/r2j5_pytest.log
tail -8 gpurun_out/r2j5_pytest.log
for c in "0.25 0.45 0" "0.001 0.6 0" "0.25 0.45 1" "0.001 0.6 1"; do set -- $c; timeout 120 python tools/run_nms.py --conf $1 --iou $2 --ml $3 --iters 10 2>&1 | tail -1; done | tee gpurun_out/r2j5_nms.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j5_nms_launches.csv \
  python tools/run_nms.py --iters 2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j5_nms_launches_c001.csv \
  python tools/run_nms.py --conf 0.001 --iou 0.6 --iters 2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j5_nms_launches_ml.csv \
  python tools/run_nms.py --conf 0.25 --iou 0.45 --ml 1 --iters 2 > /dev/null 2>&1
timeout 300 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2j5_train.log
tools/gpu_sanity.sh end
