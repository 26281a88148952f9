"""scale_boxes / clip_boxes with the reference's signatures (utils/general.py:613-626, ultralytics clip_boxes), in place on a
CUDA tensor of xyxy boxes — the step detect.py:218 / val.py:381-385 apply to the NMS output before writing results."""
from __future__ import annotations

import torch

from . import _lib
from .tensors import _stream


def _launch(boxes: torch.Tensor, pad_x: float, pad_y: float, gain: float, shape) -> torch.Tensor:
    assert boxes.is_cuda, "yolov3_b200 has no CPU path: boxes must be a CUDA tensor"
    assert boxes.dtype == torch.float32 and boxes.dim() == 2 and boxes.shape[1] >= 4 and (boxes.shape[0] == 0 or boxes.stride(1) == 1), \
        "boxes: fp32 [n, >=4] with unit inner stride (a [:, :4] view of the NMS output is fine)"
    _lib.check(_lib.lib().y3_scale_boxes(boxes.data_ptr(), boxes.shape[0], boxes.stride(0) if boxes.shape[0] else 4, float(pad_x),
                                         float(pad_y), float(gain), float(shape[1]), float(shape[0]), _stream()), "y3_scale_boxes")
    return boxes


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape, ratio_pad=None) -> torch.Tensor:
    """Rescale xyxy boxes from the letterboxed network input (img1_shape = (h, w)) to the original image (img0_shape),
    then clip to it; modifies and returns ``boxes`` like the reference."""
    if ratio_pad is None:  # calculate from img0_shape
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])  # gain = old / new
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2  # wh padding
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return _launch(boxes, pad[0], pad[1], gain, img0_shape)


def clip_boxes(boxes: torch.Tensor, shape) -> torch.Tensor:
    """Clamp xyxy boxes to an image of shape (h, w), in place."""
    return _launch(boxes, 0.0, 0.0, 1.0, shape)
