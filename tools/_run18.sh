timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x 2>&1 | tail -6 | cut -c1-300
timeout 250 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x -k oracle_autograd 2>&1 | tail -5 | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1l.json > gpurun_out/bench_r1l.log 2>&1; tail -1 gpurun_out/bench_r1l.log | cut -c1-300
Y3_CONV_XPAIR=0 timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 2 --hw 640 --time
timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 1 --hw 320 --res --time
timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 1 --hw 320 --time
timeout 100 python tools/probe_layer.py --cin 64 --cout 32 --k 1 --s 1 --hw 320 --time
timeout 100 python tools/probe_layer.py --cin 64 --cout 128 --k 3 --s 1 --hw 160 --res --time
timeout 100 python tools/probe_layer.py --cin 128 --cout 64 --k 1 --s 1 --hw 160 --time
timeout 100 python tools/probe_layer.py --cin 256 --cout 128 --k 1 --s 1 --hw 80 --time
for a in "--conf 0.25 --iou 0.45 --ml 0" "--conf 0.001 --iou 0.6 --ml 1"; do
  timeout 100 python tools/run_nms.py $a
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches_c25.csv python tools/run_nms.py --conf 0.25 --iters 1 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches_c001ml.csv python tools/run_nms.py --conf 0.001 --iou 0.6 --ml 1 --iters 1 > /dev/null 2>&1
tools/gpu_sanity.sh end
