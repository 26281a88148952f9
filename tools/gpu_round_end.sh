#!/bin/bash
# The GPU-box command sequence behind profiles/r02_*: run from the repo root on ONE B200 (e.g. through gpurun).
#   bash tools/gpu_round_end.sh        -> full GPU test suite, smoke, bench (ours + reference arm), ncu launch lists + full captures
set -u
mkdir -p gpurun_out
O=gpurun_out/r02
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -15 > ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 | tee ${O}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --per-op ${O}_per_op.json > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; tail -c 600 ${O}_bench_1gpu.json; tail -3 ${O}_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference_arm.json 2>&1; tail -c 400 ${O}_bench_reference_arm.json
# launch list of the headline command (kernel shares), DRAM traffic of every conv_tc launch of one forward
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file ${O}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --only none > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_tc --launch-skip 74 \
  --launch-count 74 --csv --log-file ${O}_ncu_conv_tc_dram.csv python tools/run_forward.py 2 > /dev/null 2>&1
# full captures: the largest conv group, wgrad, the BatchNorm backward reduction, the NMS segment kernel
timeout 400 ncu --set full --clock-control none --import-source on --source-folders yolov3_b200/csrc -k regex:conv_tc --launch-skip 84 --launch-count 1 -f \
  -o ${O}_ncu_conv_tc python tools/run_forward.py 2 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --source-folders yolov3_b200/csrc -k regex:wgrad_tc --launch-skip 140 --launch-count 2 -f \
  -o ${O}_ncu_wgrad python tools/bench_train.py --bs 8 --steps 1 --warmup 2 --no-graphs > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --source-folders yolov3_b200/csrc -k regex:bn_act_bwd --launch-skip 150 --launch-count 2 -f \
  -o ${O}_ncu_bn_bwd python tools/bench_train.py --bs 8 --steps 1 --warmup 2 --no-graphs > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --source-folders yolov3_b200/csrc -k regex:nms_ --launch-skip 12 --launch-count 6 -f \
  -o ${O}_ncu_nms python tools/run_nms.py --iters 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_train_launches.csv \
  python tools/bench_train.py --bs 8 --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
for c in "0.25 0.45 0" "0.001 0.6 0" "0.25 0.45 1"; do set -- $c; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file ${O}_nms_launches_c$1_ml$3.csv python tools/run_nms.py --conf $1 --iou $2 --ml $3 --iters 2 > /dev/null 2>&1; done
ls -la gpurun_out | grep r02 | awk '{print $5, $9}'
tools/gpu_sanity.sh end
