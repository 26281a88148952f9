"""Pipeline (detect.py:185-200 as one call): the software-pipelined ``stream()`` must return exactly what the synchronous
per-batch call returns, and both must equal forward -> non_max_suppression done by hand on the same uint8 images."""
from pathlib import Path

import numpy as np
import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu


def test_stream_equals_sync_call_and_manual_path():
    from yolov3_b200.model import Model
    from yolov3_b200.nms import non_max_suppression
    from yolov3_b200.pipeline import Pipeline

    cfg = Path(__file__).resolve().parents[1] / "yolov3_b200" / "cfg" / "yolov3-tiny.yaml"
    m = Model(cfg)
    m.load_state_dict(O.init_params(cfg, seed=0))
    n, h, w = 2, 96, 128
    pipe = Pipeline(m, n, h, w, conf_thres=0.001, iou_thres=0.6, max_det=50)
    g = torch.Generator().manual_seed(5)
    batches = [torch.randint(0, 256, (n, 3, h, w), dtype=torch.uint8, generator=g).pin_memory() for _ in range(5)]
    sync = [[d.clone() for d in pipe(b)] for b in batches]
    streamed = list(pipe.stream(batches))
    assert len(streamed) == len(batches)
    for a, b in zip(sync, streamed):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # by hand: Model.forward on the uint8 images (im/255 fused into layer 0) + the reference-signature NMS
    z, _ = m(batches[2].cuda())
    manual = non_max_suppression(z, 0.001, 0.6, max_det=50)
    for x, y in zip(manual, sync[2]):
        assert np.array_equal(x.cpu().numpy(), y.numpy())
