"""Diagnostic: per-parameter gradient error of one training step vs the CPU oracle.  python tests/diag/train_diag.py [size] [bs]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import yolo_oracle as O
from yolov3_b200.loss import ComputeLoss
from yolov3_b200.model import Model

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = ROOT / "yolov3_b200" / "cfg" / "yolov3.yaml"
params = O.init_params(cfg, seed=0)
hyp = O.scaled_hyp()
x = torch.rand(bs, 3, size, size, generator=torch.Generator().manual_seed(3))
targets = O.synth_targets(bs, seed=2)
po = {k: v.clone().requires_grad_(not ("running" in k or "anchors" in k)) for k, v in params.items()}
om = O.OracleModel(cfg, params=po, train=True)
raw_o = om.detect_raw(om.forward_features(x))
loss_o, _ = O.compute_loss(raw_o, targets, params["model.28.anchors"], hyp)
loss_o.backward()
m = Model(cfg); m.load_state_dict(params); m.hyp = hyp; m.train()
raw = m(x.cuda())
loss, _ = ComputeLoss(m)(raw, targets.cuda())
loss.backward()
torch.cuda.synchronize()
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))
print("loss", float(loss), float(loss_o), "raw", [round(rel(a.detach(), b.detach()), 5) for a, b in zip(raw, raw_o)])
P = m.device_params()
for k, v in po.items():
    if v.grad is None: continue
    g = P[k].grad
    cos = float(torch.nn.functional.cosine_similarity(g.double().cpu().flatten(), v.grad.double().flatten(), dim=0))
    print(f"{k:36s} rel {rel(g, v.grad):8.4f} cos {cos:8.5f} |ref| {float(v.grad.norm()):10.3e} |ours| {float(g.norm()):10.3e}")
