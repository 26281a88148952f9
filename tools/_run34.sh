for f in test_conv_gpu test_model_gpu test_nms_gpu test_loss_gpu test_train_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300
timeout 200 python tools/bench_train.py --bs 8 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
tools/gpu_sanity.sh end
