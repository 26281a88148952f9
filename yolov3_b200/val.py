"""``process_batch`` — the matching step of the reference's validation loop (val.py:147-188), on the device.

``process_batch(detections, labels, iouv)`` is the drop-in (one image: ``detections [N,6]``, ``labels [M,5] = (cls, xyxy)``,
returns ``bool [N, len(iouv)]`` on ``iouv.device``); ``process_batch_batched`` takes the padded NMS output of a whole batch
(``nms_batched``: ``[bs, max_det, 6]`` + counts) and the collated labels ``[nl, 6] = (image, cls, xyxy)`` and returns
``[bs, max_det, niou]`` without any device->host synchronisation — the reference copies every image's IoU matrix to the host
and runs numpy argsort/unique per threshold."""
from __future__ import annotations

import torch

from . import _lib
from .tensors import _stream

MAX_LABELS_PER_IMAGE = 1024


def process_batch_batched(det: torch.Tensor, counts: torch.Tensor | None, labels: torch.Tensor, iouv: torch.Tensor,
                          eps: float = 1e-7, overflow: torch.Tensor | None = None) -> torch.Tensor:
    assert det.is_cuda and det.dtype == torch.float32 and det.dim() == 3 and det.shape[2] == 6 and det.is_contiguous(), \
        "det: contiguous CUDA fp32 [bs, max_det, 6] (yolov3_b200 has no CPU path)"
    bs, max_det, _ = det.shape
    labels = labels.to(det.device, torch.float32).contiguous().reshape(-1, 6)
    iouv = iouv.to(det.device, torch.float32).contiguous()
    if counts is not None:
        counts = counts.to(det.device, torch.int32).contiguous()
    correct = torch.empty(bs, max_det, iouv.numel(), dtype=torch.uint8, device=det.device)
    _lib.check(_lib.lib().y3_val_match(det.data_ptr(), counts.data_ptr() if counts is not None else None, bs, max_det, max_det,
                                       labels.data_ptr() if labels.shape[0] else None, labels.shape[0], iouv.data_ptr(),
                                       iouv.numel(), float(eps), correct.data_ptr(),
                                       overflow.data_ptr() if overflow is not None else None, _stream()), "y3_val_match")
    return correct.bool()


def process_batch(detections: torch.Tensor, labels: torch.Tensor, iouv: torch.Tensor) -> torch.Tensor:
    """Drop-in for val.py:147.  detections [N,6] (x1,y1,x2,y2,conf,cls) sorted by confidence (the NMS output order — the
    result depends on it exactly as the reference's does), labels [M,5] (cls, x1,y1,x2,y2), iouv [T]."""
    assert detections.is_cuda, "yolov3_b200 has no CPU path: detections must be a CUDA tensor"
    n, m = detections.shape[0], labels.shape[0]
    if m > MAX_LABELS_PER_IMAGE:
        raise ValueError(f"process_batch: {m} labels in one image (limit {MAX_LABELS_PER_IMAGE})")
    if n == 0:
        return torch.zeros(0, iouv.numel(), dtype=torch.bool, device=iouv.device)
    det = detections.detach().float().contiguous().view(1, n, 6)
    lab = torch.cat((torch.zeros(m, 1, device=detections.device), labels.to(detections.device).float()), 1)
    return process_batch_batched(det, None, lab, iouv)[0].to(iouv.device)
