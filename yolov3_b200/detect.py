"""Detect-head decode adapter (y3_detect_decode_fwd): replaces the eval branch of Detect.forward
(reference models/yolo.py:100-110)."""
from __future__ import annotations


import torch

from . import _lib
from .tensors import _stream


def decode(raw: list[torch.Tensor], anchors_grid: torch.Tensor, stride: torch.Tensor, z: torch.Tensor | None = None):
    """raw: list of fp32 [bs,na,ny,nx,no] logits; anchors_grid [nl,na,2] in grid units (Detect.anchors); stride [nl].
    Returns z fp32 [bs, sum(na*ny*nx), no]."""
    nl = len(raw)
    bs, na, _, _, no = raw[0].shape
    levels = (_lib.DetectLevel * nl)()
    rows = 0
    anchors_px = (anchors_grid.float().cpu() * stride.float().cpu().view(-1, 1, 1))  # fp32 product like yolo.py:122
    for i, r in enumerate(raw):
        assert r.is_cuda and r.dtype == torch.float32 and r.is_contiguous()
        levels[i].raw, levels[i].ny, levels[i].nx = r.data_ptr(), r.shape[2], r.shape[3]
        levels[i].stride = float(stride[i])
        for a in range(na):
            levels[i].anchor_w[a] = float(anchors_px[i, a, 0])
            levels[i].anchor_h[a] = float(anchors_px[i, a, 1])
        rows += na * r.shape[2] * r.shape[3]
    if z is None:
        z = torch.empty(bs, rows, no, dtype=torch.float32, device=raw[0].device)
    _lib.check(_lib.lib().y3_detect_decode_fwd(levels, nl, bs, na, no, z.data_ptr(), _stream()), "y3_detect_decode_fwd")
    return z
