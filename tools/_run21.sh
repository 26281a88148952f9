timeout 900 python -m pytest tests/test_train_gpu.py tests/test_nms_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -8 | cut -c1-300
for a in "--conf 0.25 --iou 0.45 --ml 0" "--conf 0.001 --iou 0.6 --ml 0" "--conf 0.001 --iou 0.6 --ml 1"; do
  timeout 100 python tools/run_nms.py $a
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches_c25_v3.csv python tools/run_nms.py --conf 0.25 --iters 1 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches_c001ml_v3.csv python tools/run_nms.py --conf 0.001 --iou 0.6 --ml 1 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_first --launch-skip 1 --launch-count 1 -f -o gpurun_out/ncu_conv_first python tools/run_forward.py 3 > gpurun_out/ncu_conv_first.log 2>&1; tail -1 gpurun_out/ncu_conv_first.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_decode --launch-skip 1 --launch-count 1 -f -o gpurun_out/ncu_decode python tools/run_forward.py 3 > gpurun_out/ncu_decode.log 2>&1; tail -1 gpurun_out/ncu_decode.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 74 --launch-count 3 -f -o gpurun_out/ncu_early_layers python tools/run_forward.py 2 > gpurun_out/ncu_early.log 2>&1; tail -1 gpurun_out/ncu_early.log
tools/gpu_sanity.sh end
