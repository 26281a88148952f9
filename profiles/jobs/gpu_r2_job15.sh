#!/bin/bash
# round-2 GPU job 15 (N GPUs, default 2): DDP check (PDL + 4 buckets with the small tail), SyncBatchNorm, bench.py under torchrun
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2j15_${N}gpu
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/diag/check_ddp.py 2>&1 | grep -v "^W\|Warning\|warn\|^\*\|OMP_NUM" | tail -8 | tee ${O}_ddp.log
timeout 400 $TR --master-port 29512 tests/diag/check_syncbn.py 2>&1 | grep -v "^W\|Warning\|warn\|^\*\|OMP_NUM" | tail -4 | tee ${O}_syncbn.log
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > ${O}_bench.log 2> ${O}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("${O}_bench.log") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
t = d["train"]; print("train", round(t["value"], 1), round(t["ms_per_step"], 3), t["split_ms"], t["allreduce"]["buckets_mb"])
print("spp_nms", round(d["spp_nms"]["value"], 1), "nms 0.25", round(d["nms"]["conf0.25_iou0.45_single"]["input_boxes_per_s"] / 1e9, 2), "G/s")
PY
grep -v "^W\|Warning\|warn\|^\*\|OMP_NUM" ${O}_bench.err | tail -3
tools/gpu_sanity.sh end
