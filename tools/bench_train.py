"""Training-step timing (BASELINE config 4): yolov3.yaml, 640x640, bf16 storage, per-rank batch 8, coco128-shaped synthetic
targets, ComputeLoss, overlapped data-parallel gradient all-reduce over NCCL, fused SGD-nesterov + clip + EMA.  One JSON line.
    python tools/bench_train.py [--bs 8] [--steps 5] [--warmup 3]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...
bench.py runs the same step (tools/bench_workloads.py: train_step_workload) as its "train" key.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--bs", type=int, default=8)
ap.add_argument("--img", type=int, default=640)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cfg", default="yolov3.yaml")
ap.add_argument("--no-graphs", action="store_true", help="eager launches (ncu launch lists)")
ap.add_argument("--torch-optim", action="store_true", help="torch.optim.SGD on the parameter views instead of the fused step")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
sys.path.insert(0, str(ROOT / "tools"))
from bench_workloads import train_step_workload  # noqa: E402

res = train_step_workload(dev, rank, world, bs=a.bs, img=a.img, steps=a.steps, warmup=a.warmup, cfg=a.cfg,
                          use_graphs=not a.no_graphs, torch_optim=a.torch_optim)
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
