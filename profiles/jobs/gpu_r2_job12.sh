#!/bin/bash
# round-2 GPU job 12: whole-tile residual prefetch (conv_tc N = 256) + BatchNorm look-ahead kernels: GPU tests, bench, per-op
set -u
mkdir -p gpurun_out
O=gpurun_out/r2j12
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -8 > ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --only train --per-op ${O}_per_op.json > ${O}_bench.json 2> ${O}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("${O}_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4))
t = d["train"]; print("train", round(t["value"], 1), round(t["ms_per_step"], 3), t["split_ms"])
ops = json.load(open("${O}_per_op.json"))
for o in ops:
    if "+res" in o["shape"] and "@80x80" in o["shape"]: print("  ", o["shape"], round(o["ms"] * 1e3, 1), "us", round(o["tflops"], 1), "TF")
PY
tail -2 ${O}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_train_launches.csv \
  python tools/bench_train.py --bs 8 --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
python tools/launch_summary.py ${O}_train_launches.csv --last-step sgd_step | head -14
tools/gpu_sanity.sh end
