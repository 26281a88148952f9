"""ncu / timing target for the device NMS: BASELINE config 5 (bs 32 x 25200 x 85 synthetic predictions, seed 3).
  python tools/run_nms.py [--conf 0.25] [--iou 0.45] [--ml 0] [--iters 3]"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conf", type=float, default=0.25)
    ap.add_argument("--iou", type=float, default=0.45)
    ap.add_argument("--ml", type=int, default=0)
    ap.add_argument("--bs", type=int, default=32)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    import torch
    from yolov3_b200 import synth as O
    from yolov3_b200.nms import nms_batched

    pred = O.synth_predictions(a.bs, n_rows=25200, nc=80, seed=3).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(a.iters + 1):
        if it == 1:
            ev[0].record()
        out, counts, overflow, _ = nms_batched(pred, a.conf, a.iou, multi_label=bool(a.ml), max_det=300)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / a.iters
    print(f"nms conf {a.conf} iou {a.iou} ml {a.ml}: {ms:.3f} ms/batch, {a.bs * 25200 / ms / 1e6:.3f} G boxes/s, "
          f"kept {int(counts.sum())}, overflow {int(overflow.max())}")


if __name__ == "__main__":
    main()
