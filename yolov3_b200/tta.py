"""Test-time augmentation and model ensembling with the reference's semantics:

``forward_augment(model, x)`` = ``Model._forward_augment`` (models/yolo.py:239-280): three views of the batch (scales 1, 0.83,
0.67; the 0.83 view flipped left-right) built by ``scale_img``, one engine per view shape, the decoded rows de-scaled / de-flipped,
the P5 rows of the full-size view and the P3 rows of the smallest view dropped, everything concatenated: ``(z_aug, None)``.
``Ensemble`` = models/experimental.py:74-85 (NMS ensemble: concatenate every member's rows); ``attempt_load`` = :88-136."""
from __future__ import annotations

import math

import torch

from . import _lib
from .tensors import _stream

SCALES = (1, 0.83, 0.67)   # models/yolo.py:242
FLIPS = (None, 3, None)    # :243 (3 = left-right)


def scale_img(img: torch.Tensor, ratio=1.0, same_shape=False, gs=32, flip_lr=False) -> torch.Tensor:
    """ultralytics scale_img (used at models/yolo.py:246) on a CUDA fp32 NCHW batch; ``flip_lr`` folds the reference's
    ``x.flip(3)`` into the same pass."""
    assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 4, "scale_img: CUDA fp32 [n,c,h,w]"
    img = img.contiguous()
    n, c, h, w = img.shape
    if ratio == 1.0 and not flip_lr:
        return img
    rh, rw = (h, w) if ratio == 1.0 else (int(h * ratio), int(w * ratio))
    oh, ow = (rh, rw) if (same_shape or ratio == 1.0) else tuple(math.ceil(v * ratio / gs) * gs for v in (h, w))
    out = torch.empty(n, c, oh, ow, dtype=torch.float32, device=img.device)
    _lib.check(_lib.lib().y3_scale_img_f32(img.data_ptr(), n, c, h, w, rh, rw, oh, ow, int(bool(flip_lr)), 0.447, out.data_ptr(),
                                           _stream()), "y3_scale_img_f32")
    return out


def clip_rows(rows_first: int, rows_last: int, nl: int):
    """_clip_augmented (models/yolo.py:268-278): rows dropped from the END of the first view and the START of the last one."""
    g = sum(4 ** x for x in range(nl))
    e = 1
    drop_first = (rows_first // g) * sum(4 ** x for x in range(e))
    drop_last = (rows_last // g) * sum(4 ** (nl - 1 - x) for x in range(e))
    return drop_first, drop_last


def forward_augment(model, x: torch.Tensor):
    """Returns (z_aug [bs, rows, no], None) like the reference's augmented inference."""
    if x.dtype == torch.uint8:
        x = x.float() / 255  # TTA resamples the image: it works on the float image like the reference (detect.py:187-191)
    x = x.float().contiguous()
    n, _, h, w = x.shape
    gs = int(model.stride.max())
    det = model.detect
    # rows every view will produce, known from its shape alone (views may share an engine — 0.83 x 96 pads back to 96 — so each
    # view is merged into the output right after it ran, before the engine's z buffer is reused)
    shapes = []
    for si in SCALES:
        vh, vw = (h, w) if si == 1 else tuple(math.ceil(v * si / gs) * gs for v in (h, w))
        shapes.append((vh, vw, det.na * sum((vh // int(s)) * (vw // int(s)) for s in det.stride.tolist())))
    drop_first, drop_last = clip_rows(shapes[0][2], shapes[-1][2], det.nl)
    ranges = [(0, r) for _, _, r in shapes]
    ranges[0] = (0, ranges[0][1] - drop_first)
    ranges[-1] = (drop_last, ranges[-1][1])
    total = sum(b - a for a, b in ranges)
    out = torch.empty(n, total, det.no, dtype=torch.float32, device=x.device)
    off = 0
    L = _lib.lib()
    for si, fi, (vh, vw, rows), (a, b) in zip(SCALES, FLIPS, shapes, ranges):
        xi = scale_img(x, si, gs=gs, flip_lr=fi == 3)
        assert tuple(xi.shape[2:]) == (vh, vw)
        e = model.engine(n, vh, vw, torch.float32)
        z, _ = e.run(xi)
        assert z.shape[1] == rows
        _lib.check(L.y3_tta_merge(z.data_ptr(), n, rows, det.no, a, b, float(si), int(fi == 3), float(w), out.data_ptr(), total, off,
                                  _stream()), "y3_tta_merge")
        off += b - a
    return out, None


class Ensemble(list):
    """models/experimental.py:74-85: ``forward`` concatenates the members' inference rows (NMS ensemble)."""

    def forward(self, x, augment=False, profile=False, visualize=False):
        y = [m(x, augment=augment)[0] for m in self]
        return torch.cat(y, 1), None

    __call__ = forward

    def eval(self):
        for m in self:
            m.eval()
        return self


def attempt_load(weights, device=None, inplace=True, fuse=True):
    """models/experimental.py:88-136: one checkpoint -> ``Model``; several -> ``Ensemble`` carrying names / nc / yaml of the
    first member and the largest stride."""
    from .backend import _load

    device = torch.device(device if device is not None else "cuda")
    members = Ensemble()
    for w in weights if isinstance(weights, (list, tuple)) else [weights]:
        m = _load(w, device)
        m.inplace = inplace
        members.append(m.fuse().eval() if fuse else m.eval())
    if len(members) == 1:
        return members[-1]
    for k in ("names", "nc", "yaml"):
        setattr(members, k, getattr(members[0], k))
    members.stride = members[int(torch.argmax(torch.tensor([float(m.stride.max()) for m in members])))].stride
    assert all(members[0].nc == m.nc for m in members), f"Models have different class counts: {[m.nc for m in members]}"
    return members
