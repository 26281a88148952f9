#!/bin/bash
# round-2 GPU job 18 (2 GPUs): does capping NCCL's CTAs reduce the interference of the overlapped all-reduce with the backward?
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ctas in default 8 4; do
  if [ $ctas = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --only train > gpurun_out/r2j18_$ctas.log 2> gpurun_out/r2j18_$ctas.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2j18_$ctas.log") if l.startswith("{")][-1])
    t = d["train"]; print("NCCL_MAX_CTAS=$ctas train", round(t["value"], 1), round(t["ms_per_step"], 3), t["split_ms"])
except Exception as e:
    print("NCCL_MAX_CTAS=$ctas failed", e)
PY
done
