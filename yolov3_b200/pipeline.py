"""``Pipeline`` — the detect.py inner loop as one call (reference detect.py:185-200): host uint8 images ->
device (``im.to(device)``) -> ``im.float()/255`` + Model.forward (fused into layer 0) -> Detect decode ->
non_max_suppression -> detections back on the host.  Everything between the H2D copy of the images and the D2H copy
of the kept boxes stays on the device with no host synchronisation.

``Pipeline(images)`` is the synchronous per-batch call of the reference loop.  ``Pipeline.stream(batches)`` runs the same
steps software-pipelined over two buffers: the H2D copy of batch i+1 (copy stream) and the host-side read-out of batch
i-1 overlap the forward + NMS of batch i, results come back in order."""
from __future__ import annotations

import torch

from .model import Model
from .nms import nms_batched


class Pipeline:
    def __init__(self, model: Model, n, h, w, conf_thres=0.25, iou_thres=0.45, max_det=300, multi_label=False,
                 agnostic=False, classes=None, use_graph=True):
        self.model = model
        self.shape = (n, h, w)
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
        self._stream_state = None

    def _fresh_engine(self):
        """The engine this pipeline launches; re-lowered (and re-captured) when the model's weights changed since it was
        built (load_state_dict, a training phase followed by eval()): an Engine never runs on stale packed weights."""
        if self.engine.stale:
            self.engine = self.model.engine(*self.shape, torch.uint8, 255.0)
            if self.use_graph and self.engine.graph is None:
                self.engine.capture()
        return self.engine

    def __call__(self, images_u8: torch.Tensor):
        """images_u8: HOST uint8 [n,3,h,w] (pinned for an async copy).  Returns list of host tensors [k,6] (copies: the
        pinned read-back buffer is reused by the next call)."""
        e = self._fresh_engine()
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
        return [self.host_out[i, : int(self.host_cnt[0, i])].clone() for i in range(self.host_out.shape[0])]

    # ------------------------------------------------------------------------------------------------ streaming
    def _slots(self):
        if self._stream_state is None:
            e = self.engine
            n, max_det = self.host_out.shape[0], self.host_out.shape[1]
            self._stream_state = dict(
                copy=torch.cuda.Stream(device=self.dev),
                slots=[dict(dev_in=torch.empty_like(e.static_in),
                            host_out=torch.empty(n, max_det, 6, dtype=torch.float32).pin_memory(),
                            host_cnt=torch.empty(2, n, dtype=torch.int32).pin_memory(),
                            h2d=torch.cuda.Event(), free=torch.cuda.Event(), done=torch.cuda.Event()) for _ in range(2)])
            for sl in self._stream_state["slots"]:
                sl["free"].record()
        return self._stream_state

    def _collect(self, sl):
        sl["done"].synchronize()
        if int(sl["host_cnt"][1].max()):
            raise RuntimeError("NMS candidate capacity exceeded; rerun through non_max_suppression() for the exact retry")
        return [sl["host_out"][i, : int(sl["host_cnt"][0, i])].clone() for i in range(sl["host_out"].shape[0])]

    def stream(self, batches):
        """Generator over per-batch detections (same values as ``self(batch)``) for an iterable of HOST uint8 batches
        (pinned for a truly asynchronous copy)."""
        st = self._slots()
        e, cs = self._fresh_engine(), st["copy"]
        main = torch.cuda.current_stream()
        pending = None
        for i, images_u8 in enumerate(batches):
            sl = st["slots"][i & 1]
            with torch.cuda.stream(cs):
                cs.wait_event(sl["free"])  # the forward that read this slot's device buffer two steps ago has consumed it
                sl["dev_in"].copy_(images_u8, non_blocking=True)
                sl["h2d"].record(cs)
            main.wait_event(sl["h2d"])
            e.static_in.copy_(sl["dev_in"], non_blocking=True)  # device-to-device, 39 MB at bs 32: ~15 us
            sl["free"].record(main)
            if self.use_graph:
                e.replay()
            else:
                e.run(None)
            out, counts, overflow, _ = nms_batched(e.z, **self.kw)
            sl["host_out"].copy_(out, non_blocking=True)
            sl["host_cnt"][0].copy_(counts, non_blocking=True)
            sl["host_cnt"][1].copy_(overflow, non_blocking=True)
            sl["done"].record(main)
            if pending is not None:
                yield self._collect(pending)
            pending = sl
        if pending is not None:
            yield self._collect(pending)
