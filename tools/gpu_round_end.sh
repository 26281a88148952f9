#!/bin/bash
# The GPU-box command sequence behind profiles/r01_*: run from the repo root on a B200 (e.g. through gpurun).
#   bash tools/gpu_round_end.sh            -> tests, smoke, bench (ours + reference arm), ncu launch list + full capture
set -u
mkdir -p gpurun_out
for f in test_conv_gpu test_model_gpu test_nms_gpu test_loss_gpu test_pipeline_gpu test_train_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2
done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --per-op gpurun_out/per_op.json > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1
# launch list of the bench command (shares) and one full capture of the largest layer group (128->256 3x3 + residual)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 84 --launch-count 1 -f \
  -o gpurun_out/ncu_conv_tc python tools/run_forward.py 2 > /dev/null 2>&1
# 2 GPUs (gpurun --gpus 2): python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#   bench.py --gpus 2 ; ... tools/bench_train.py ; ... tests/diag/check_syncbn.py
tools/gpu_sanity.sh end
