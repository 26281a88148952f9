timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3 | cut -c1-300
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | grep -E "^E  |passed|failed" | head -12 | cut -c1-1500
tools/gpu_sanity.sh end
