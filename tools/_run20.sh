timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider -x 2>&1 | tail -8 | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_first --launch-skip 2 --launch-count 1 -f -o gpurun_out/ncu_conv_first python tools/run_forward.py > gpurun_out/ncu_conv_first.log 2>&1; tail -2 gpurun_out/ncu_conv_first.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:head_decode --launch-skip 2 --launch-count 1 -f -o gpurun_out/ncu_decode python tools/run_forward.py > gpurun_out/ncu_decode.log 2>&1; tail -2 gpurun_out/ncu_decode.log
tools/gpu_sanity.sh end
