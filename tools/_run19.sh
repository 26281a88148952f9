timeout 300 python -m pytest tests/test_nms_gpu.py tests/test_model_gpu.py -m gpu -q --timeout 200 -p no:cacheprovider -x 2>&1 | tail -6 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-op gpurun_out/per_op_r1m.json > gpurun_out/bench_r1m.log 2>&1; tail -1 gpurun_out/bench_r1m.log | cut -c1-200; tail -1 gpurun_out/bench_r1m.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']); print(d.get('nms'))"
for a in "--conf 0.25 --iou 0.45 --ml 0" "--conf 0.001 --iou 0.6 --ml 1"; do
  timeout 100 python tools/run_nms.py $a
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches_c25_v2.csv python tools/run_nms.py --conf 0.25 --iters 1 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nms_launches_c001ml_v2.csv python tools/run_nms.py --conf 0.001 --iou 0.6 --ml 1 --iters 1 > /dev/null 2>&1
tools/gpu_sanity.sh end
