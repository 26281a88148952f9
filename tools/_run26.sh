TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; tail -1 gpurun_out/bench_2gpu.log | cut -c1-400
timeout 300 $TR tools/bench_train.py --bs 8 --steps 5 --warmup 3 > gpurun_out/train_2gpu.log 2>&1; tail -1 gpurun_out/train_2gpu.log | cut -c1-700
timeout 300 $TR tools/check_syncbn.py > gpurun_out/syncbn_2gpu.log 2>&1; grep -E "sync-bn|SYNCBN|Error|error" gpurun_out/syncbn_2gpu.log | head -5 | cut -c1-300
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-700
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1q.json > gpurun_out/bench_r1q.log 2>&1; tail -1 gpurun_out/bench_r1q.log | cut -c1-200
nvidia-smi --query-gpu=name,temperature.gpu --format=csv,noheader
