#!/bin/bash
# round-2 GPU job 17: two-phase suppression-matrix kernel for segments up to 512 members: NMS parity tests, sweep, launch lists
set -u
mkdir -p gpurun_out
O=gpurun_out/r2j17
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "nms or pipeline or seam or val or tta" 2>&1 | tail -6 > ${O}_pytest.log; tail -3 ${O}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --only nms > ${O}_bench.json 2> ${O}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("${O}_bench.json") if l.startswith("{")][-1])
for k, v in d["nms"].items():
    if isinstance(v, dict): print(k, round(v["ms_per_batch"], 4), "ms", round(v["input_boxes_per_s"] / 1e9, 3), "G boxes/s  kept", round(v["kept_per_image"], 1))
PY
tail -2 ${O}_bench.err
for c in "0.001 0.6 0" "0.25 0.45 1" "0.001 0.6 1"; do set -- $c; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file ${O}_nms_launches_c$1_ml$3.csv python tools/run_nms.py --conf $1 --iou $2 --ml $3 --iters 2 > /dev/null 2>&1
  python tools/launch_summary.py ${O}_nms_launches_c$1_ml$3.csv --from-last nms_candidates --title "conf $1 ml $3" | head -10 || true; done
tools/gpu_sanity.sh end
