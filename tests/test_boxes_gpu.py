"""scale_boxes / clip_boxes (utils/general.py:613-626) on the device: bit-exact against fixtures produced by the reference."""
import ast
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"


def test_scale_boxes_golden_bit_exact():
    from yolov3_b200.boxes import clip_boxes, scale_boxes

    g = np.load(G / "scale_boxes_cases.npz")
    for ci in range(sum(k.startswith("geom") for k in g.files)):
        s1, s0, rp = ast.literal_eval(str(g[f"geom{ci}"]))
        det = torch.from_numpy(g[f"in{ci}"]).cuda()  # [200, 6] like the NMS output: the boxes are the [:, :4] view
        ret = scale_boxes(s1, det[:, :4], s0, rp)
        assert ret.data_ptr() == det.data_ptr()  # in place, returns its argument like the reference
        assert np.array_equal(det.cpu().numpy(), g[f"out{ci}"])  # columns 4, 5 untouched
    b = torch.tensor([[-5.0, 10.0, 700.0, 500.0], [float("nan"), 1.0, 2.0, 3.0]], device="cuda")
    clip_boxes(b, (480, 640))
    assert b[0].tolist() == [0.0, 10.0, 640.0, 480.0] and torch.isnan(b[1, 0]) and b[1, 1:].tolist() == [1.0, 2.0, 3.0]
    assert scale_boxes((640, 640), torch.zeros(0, 4, device="cuda"), (480, 640)).shape == (0, 4)
    with pytest.raises(AssertionError):
        scale_boxes((640, 640), torch.zeros(3, 4), (480, 640))  # no CPU path
