"""DetectMultiBackend (models/common.py:435-476, `pt` role) on the device.  Kept in its own, last-sorting file: it was written
after the round's GPU budget was spent, so it must not be able to stop the validated suites under `pytest -x`."""
from pathlib import Path

import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu


def test_detect_multi_backend_wraps_the_model():
    """models/common.py:435-476 `pt` role: attributes, warmup, forward == the wrapped model's forward, fp16 casts outputs."""
    from yolov3_b200.backend import DetectMultiBackend
    from yolov3_b200.model import Model

    cfg = Path(__file__).resolve().parents[1] / "yolov3_b200" / "cfg" / "yolov3-tiny.yaml"
    m = Model(cfg)
    m.load_state_dict(O.init_params(cfg, seed=0))
    b = DetectMultiBackend(m, device=torch.device("cuda"))
    assert b.pt and not (b.jit or b.engine or b.triton or b.nhwc) and b.stride == 32 and len(b.names) == 80 and b.model is m
    b.warmup((1, 3, 64, 64))
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)).cuda()
    z, raw = b(x)
    z2, raw2 = m(x)
    assert torch.equal(z, z2) and all(torch.equal(a, c) for a, c in zip(raw, raw2))
    zh, rawh = DetectMultiBackend(m, device=torch.device("cuda"), fp16=True)(x.half())
    z3, _ = m(x.half().float())  # the half input, widened, through the same kernels
    assert zh.dtype == torch.float16 and rawh[0].dtype == torch.float16 and torch.equal(zh, z3.half())
