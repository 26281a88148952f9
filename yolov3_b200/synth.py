"""Synthetic workloads of the BASELINE configs — inputs for bench.py and the tools, so that nothing outside
``tests/`` / ``smoke()`` / the CPU-baseline leg needs ``oracle/``.  The oracle carries its own copies for the tests;
``tests/test_host_cpu.py`` checks that both generate identical tensors."""
from __future__ import annotations

import math

import torch

# data/hyps/hyp.scratch-low.yaml values on the loss path (the keys ComputeLoss reads, utils/loss.py:104-129)
DEFAULT_HYP = dict(box=0.05, obj=1.0, cls=0.5, cls_pw=1.0, obj_pw=1.0, fl_gamma=0.0, anchor_t=4.0, label_smoothing=0.0)


def scaled_hyp(hyp=None, nl=3, nc=80, imgsz=640):
    """train.py:326-329: box *= 3/nl, cls *= nc/80 * 3/nl, obj *= (imgsz/640)^2 * 3/nl."""
    h = dict(DEFAULT_HYP if hyp is None else hyp)
    h["box"] *= 3 / nl
    h["cls"] *= nc / 80 * 3 / nl
    h["obj"] *= (imgsz / 640) ** 2 * 3 / nl
    return h


def synth_predictions(bs, n_rows=25200, nc=80, seed=3, imgsz=640):
    """Config 5 NMS input [bs, n_rows, 5+nc] fp32: xy~U(0,imgsz), wh~U(4,204), obj~U(0,1)^6, cls~U(0,1)^4."""
    g = torch.Generator().manual_seed(seed)
    p = torch.empty(bs, n_rows, 5 + nc)
    p[..., 0:2] = torch.rand(bs, n_rows, 2, generator=g) * imgsz
    p[..., 2:4] = torch.rand(bs, n_rows, 2, generator=g) * 200 + 4
    p[..., 4] = torch.rand(bs, n_rows, generator=g) ** 6
    p[..., 5:] = torch.rand(bs, n_rows, nc, generator=g) ** 4
    return p


def synth_targets(bs, nc=80, seed=2):
    """Config 4 targets, coco128-shaped: n~Poisson(7.3) clipped to [1,40] per image, cls~U{0..nc-1}, xy~U(.05,.95),
    wh~LogUniform(.02,.6) clipped inside the image; layout [nt,6] = (img, cls, x, y, w, h) as collate_fn builds it
    (utils/dataloaders.py:825-830)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(bs):
        n = int(torch.poisson(torch.tensor(7.3), generator=g).clamp(1, 40))
        cls = torch.randint(0, nc, (n,), generator=g).float()
        xy = torch.rand(n, 2, generator=g) * 0.9 + 0.05
        wh = torch.exp(torch.rand(n, 2, generator=g) * (math.log(0.6) - math.log(0.02)) + math.log(0.02))
        wh = torch.minimum(wh, 2 * torch.minimum(xy, 1 - xy))
        rows.append(torch.cat((torch.full((n, 1), float(b)), cls[:, None], xy, wh), 1))
    return torch.cat(rows, 0)
