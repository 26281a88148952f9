timeout 200 python tools/train_repeat.py --graphs 0 2>&1 | tail -5 | cut -c1-500
timeout 200 python tools/train_repeat.py --graphs 1 2>&1 | tail -5 | cut -c1-500
timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 2 --hw 640 --time
timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 1 --hw 320 --res --time
timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 1 --hw 320 --time
Y3_CONV_STAGED=0 timeout 100 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 1 --hw 320 --res --time
timeout 100 python tools/probe_layer.py --cin 64 --cout 32 --k 1 --s 1 --hw 320 --time
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 3 --launch-count 1 -f -o gpurun_out/ncu_layer1 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 2 --hw 640 > gpurun_out/ncu_layer1.log 2>&1; tail -2 gpurun_out/ncu_layer1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 3 --launch-count 1 -f -o gpurun_out/ncu_layer3 python tools/probe_layer.py --cin 32 --cout 64 --k 3 --s 1 --hw 320 --res > gpurun_out/ncu_layer3.log 2>&1; tail -2 gpurun_out/ncu_layer3.log
timeout 250 python -m pytest tests/test_nms_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x 2>&1 | tail -3
tools/gpu_sanity.sh end
