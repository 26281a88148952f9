"""``Pipeline`` — the detect.py inner loop as one call (reference detect.py:185-200): host uint8 images ->
device (``im.to(device)``) -> ``im.float()/255`` + Model.forward (fused into layer 0) -> Detect decode ->
non_max_suppression -> detections back on the host.  Everything between the H2D copy of the images and the D2H copy
of the kept boxes stays on the device with no host synchronisation."""
from __future__ import annotations

import torch

from .model import Model
from .nms import nms_batched


class Pipeline:
    def __init__(self, model: Model, n, h, w, conf_thres=0.25, iou_thres=0.45, max_det=300, multi_label=False,
                 agnostic=False, classes=None, use_graph=True):
        self.model = model
        self.engine = model.engine(n, h, w, torch.uint8, 255.0)
        self.kw = dict(conf_thres=conf_thres, iou_thres=iou_thres, max_det=max_det, multi_label=multi_label,
                       agnostic=agnostic, classes=classes)
        dev = model.device
        self.host_out = torch.empty(n, max_det, 6, dtype=torch.float32).pin_memory()
        self.host_cnt = torch.empty(2, n, dtype=torch.int32).pin_memory()
        self.h2d_bytes = n * model.ch * h * w  # uint8
        self.d2h_bytes = self.host_out.numel() * 4 + self.host_cnt.numel() * 4
        self.use_graph = use_graph
        if use_graph:
            self.engine.capture()
        self.dev = dev

    def __call__(self, images_u8: torch.Tensor):
        """images_u8: HOST uint8 [n,3,h,w] (pinned for an async copy).  Returns list of host tensors [k,6]."""
        e = self.engine
        e.static_in.copy_(images_u8, non_blocking=True)
        if self.use_graph:
            e.replay()
        else:
            e.run(None)
        out, counts, overflow, _ = nms_batched(e.z, **self.kw)
        self.host_out.copy_(out, non_blocking=True)
        self.host_cnt[0].copy_(counts, non_blocking=True)
        self.host_cnt[1].copy_(overflow, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if int(self.host_cnt[1].max()):
            raise RuntimeError("NMS candidate capacity exceeded; rerun through non_max_suppression() for the exact retry")
        return [self.host_out[i, : int(self.host_cnt[0, i])] for i in range(self.host_out.shape[0])]
