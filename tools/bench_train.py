"""Training-step timing (BASELINE config 4): yolov3.yaml, 640x640, bf16 storage, per-rank batch 8, coco128-shaped synthetic
targets, ComputeLoss, data-parallel gradient all-reduce over NCCL.  Prints one JSON line on rank 0.
    python tools/bench_train.py [--bs 8] [--steps 5] [--warmup 2]            (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...
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
from yolov3_b200 import parallel  # noqa: E402
from yolov3_b200 import synth as O  # noqa: E402  (synthetic targets / hyp)
from yolov3_b200.loss import ComputeLoss  # noqa: E402
from yolov3_b200.model import Model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bs", type=int, default=8)
ap.add_argument("--img", type=int, default=640)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--no-graphs", action="store_true", help="eager launches (ncu launch lists)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
if a.no_graphs:
    from yolov3_b200.train import TrainEngine

    TrainEngine.use_graphs = False
torch.manual_seed(0)
m = Model("yolov3.yaml", device=dev)
m.hyp = O.scaled_hyp()
m.train()
params = list(m.device_params().values())
parallel.broadcast_parameters([p for p in params], 0)
opt = torch.optim.SGD(list(m.parameters()), lr=1e-3, momentum=0.937, nesterov=True)
loss_fn = ComputeLoss(m)
x = torch.rand(a.bs, 3, a.img, a.img, device=dev)
targets = O.synth_targets(a.bs, seed=2 + rank).to(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = [0.0] * 5
flat = None
for it in range(a.warmup + a.steps):
    ev[0].record()
    pred = m(x)
    ev[1].record()
    loss, items = loss_fn(pred, targets)
    loss = parallel.scale_loss(loss)
    ev[2].record()
    loss.backward()
    ev[3].record()
    flat = parallel.allreduce_gradients(m.parameters(), flat=flat)
    ev[4].record()
    opt.step()
    opt.zero_grad(set_to_none=False)
    ev[5].record()
    torch.cuda.synchronize()
    if it >= a.warmup:
        for k in range(5):
            acc[k] += ev[k].elapsed_time(ev[k + 1])
ms = [v / a.steps for v in acc]
tot = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
if rank == 0:
    total = float(tot.item())
    print(json.dumps({"metric": "train images/sec @640 YOLOv3 (fwd+loss+bwd+allreduce+SGD)", "value": world * a.bs / (total / 1e3),
                      "unit": "images/s", "n_gpus": world, "batch_per_gpu": a.bs, "ms_per_step": total,
                      "split_ms": dict(zip(["forward", "loss_fwd", "backward", "allreduce", "optimizer"], [round(v, 3) for v in ms])),
                      "loss": float(loss.detach()), "loss_items": [float(v) for v in items],
                      "grad_bytes_allreduced": sum(p.numel() for p in m.parameters()) * 4, "dtype": "bf16 storage / fp32 master"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
