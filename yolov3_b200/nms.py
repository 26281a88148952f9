"""non_max_suppression — same call signature, assertions and return type as the reference (utils/general.py:630-750),
computed by the sync-free device pipeline in csrc/y3_nms.cu.  ``nms_batched`` is the sync-free form (padded outputs
plus per-image counts, everything stays on the device); ``non_max_suppression`` adds the one device->host read the
reference's list-of-tensors return type forces."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .tensors import _stream

MAX_NMS = 30000  # utils/general.py:674
MAX_WH = 7680.0  # utils/general.py:673
_ws_cache: dict = {}


def _workspace(bs, cap, device):
    key = (bs, cap, str(device))
    ws = _ws_cache.get(key)
    if ws is None:
        nbytes = _lib.lib().y3_nms_workspace_bytes(bs, cap)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache.clear()  # keep a single workspace alive
        _ws_cache[key] = ws
    return ws


def nms_batched(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                multi_label=False, max_det=300, cap=None, want_src=False):
    """Device-side batched NMS.  Returns (out[bs,max_det,6], counts[bs] int32, overflow[bs] int32, src or None)."""
    assert prediction.is_cuda, "yolov3_b200 has no CPU path: prediction must be a CUDA tensor"
    pred = prediction.detach()
    if pred.dtype != torch.float32 or not pred.is_contiguous():
        pred = pred.float().contiguous()
    bs, n_rows, no = pred.shape
    nc = no - 5
    L = _lib.lib()
    ml = bool(multi_label) and nc > 1
    if cap is None:
        cap = L.y3_nms_default_capacity(n_rows, nc, int(ml))
    ws = _workspace(bs, cap, pred.device)
    out = torch.empty(bs, max_det, 6, dtype=torch.float32, device=pred.device)
    counts = torch.empty(bs, dtype=torch.int32, device=pred.device)
    overflow = torch.empty(bs, dtype=torch.int32, device=pred.device)
    src = torch.empty(bs, max_det, 2, dtype=torch.int32, device=pred.device) if want_src else None
    p = _lib.NmsParams()
    p.bs, p.n_rows, p.nc = bs, n_rows, nc
    p.conf_thres, p.iou_thres = float(conf_thres), float(iou_thres)
    p.multi_label, p.agnostic = int(ml), int(bool(agnostic))
    p.max_det, p.max_nms, p.max_wh, p.cap = int(max_det), MAX_NMS, MAX_WH, int(cap)
    if classes is not None:
        arr = (C.c_int32 * len(classes))(*[int(c) for c in classes])
        p.classes, p.n_classes = arr, len(classes)
    _lib.check(L.y3_nms_batched(pred.data_ptr(), C.byref(p), ws.data_ptr(), ws.numel(), out.data_ptr(),
                                src.data_ptr() if src is not None else None, counts.data_ptr(), overflow.data_ptr(),
                                _stream()), "y3_nms_batched")
    return out, counts, overflow, src


def _append_labels(pred: torch.Tensor, labels):
    """Autolabel priors (utils/general.py:689-695): each image's (cls, x, y, w, h) rows become extra candidates with
    obj = 1 and a one-hot class, placed after that image's predictions (the reference concatenates them after its
    confidence filter, so candidate order is identical).  Images with fewer labels get obj = 0 filler rows, which the
    confidence filter drops."""
    bs, _, no = pred.shape
    lmax = max(len(l) for l in labels)
    v = torch.zeros(bs, lmax, no, dtype=pred.dtype, device=pred.device)
    for xi, lb in enumerate(labels):
        if len(lb):
            lb = torch.as_tensor(lb, dtype=torch.float32, device=pred.device).reshape(-1, 5)
            m = lb.shape[0]
            v[xi, :m, :4] = lb[:, 1:5].to(pred.dtype)
            v[xi, :m, 4] = 1.0
            v[xi, torch.arange(m, device=pred.device), lb[:, 0].long() + 5] = 1.0
    return torch.cat((pred, v), 1)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nm=0, return_src=False):
    """Drop-in for utils/general.py:630.  Returns list[Tensor[n,6]] (xyxy, conf, cls), rows sorted by conf desc."""
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):  # (inference_out, loss_out), general.py:660-661
        prediction = prediction[0]
    if nm:
        raise NotImplementedError("mask coefficients (nm>0) are not part of the YOLOv3 detection path")
    if labels and any(len(l) for l in labels):
        prediction = _append_labels(prediction, labels)
    cap = None
    while True:
        out, counts, overflow, src = nms_batched(prediction, conf_thres, iou_thres, classes, agnostic, multi_label,
                                                 max_det, cap, want_src=return_src)
        host = torch.stack((counts, overflow)).cpu()  # the single device->host read
        worst = int(host[1].max())
        if worst == 0:
            break
        cap = 1 << (worst - 1).bit_length()  # exact retry: every candidate fits
    res = [out[i, : int(host[0, i])] for i in range(out.shape[0])]
    if return_src:
        return res, [src[i, : int(host[0, i])] for i in range(out.shape[0])]
    return res
