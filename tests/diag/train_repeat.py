"""Diagnostic: run the same training step several times (same images, same targets, no optimizer step) and report how far
the parameter gradients of step k are from those of step 1 — eager launches first, then through the captured CUDA graphs.
  python tests/diag/train_repeat.py [--graphs 0|1] [--steps 4]"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    import torch
    import yolo_oracle as O

    from yolov3_b200.loss import ComputeLoss
    from yolov3_b200.model import Model
    from yolov3_b200.train import TrainEngine

    TrainEngine.use_graphs = bool(a.graphs)
    cfg = ROOT / "yolov3_b200" / "cfg" / "yolov3.yaml"
    m = Model(cfg)
    m.load_state_dict(O.init_params(cfg, seed=0))
    m.hyp = O.scaled_hyp()
    m.train()
    x = torch.rand(4, 3, 96, 96, generator=torch.Generator().manual_seed(3)).cuda()
    t = O.synth_targets(4, seed=2).cuda()
    P = m.device_params()
    names = [k for k, v in P.items() if v.requires_grad]
    first = None
    for step in range(a.steps):
        for k in names:
            P[k].grad = None
        raw = m(x)
        loss, _ = ComputeLoss(m)(raw, t)
        loss.backward()
        torch.cuda.synchronize()
        g = {k: P[k].grad.clone() for k in names if P[k].grad is not None}
        raws = [r.detach().clone() for r in raw]
        if first is None:
            first, first_raw = g, raws
            print(f"step 0 loss {float(loss):.6f}")
            continue
        errs = sorted(((rel_l2(g[k], first[k]), k) for k in g), reverse=True)
        fw = max(rel_l2(r, r0) for r, r0 in zip(raws, first_raw))
        print(f"step {step} loss {float(loss):.6f}  forward raw rel-L2 {fw:.2e}  worst grads: "
              + ", ".join(f"{k}={e:.3f}" for e, k in errs[:4]) + f"  median {errs[len(errs) // 2][0]:.2e}")


if __name__ == "__main__":
    main()
