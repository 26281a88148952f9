"""Device NMS (y3_nms_batched) against the reference goldens (bit-exact rows and kept (row, class) sets) and against
the CPU oracle on the full-size BASELINE config-5 workload."""
import ast
from pathlib import Path

import numpy as np
import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"


def _cases():
    g = np.load(G / "nms_cases.npz")
    return sorted({k.split("/")[0] for k in g.files if "/" in k})


@pytest.mark.parametrize("case", _cases())
def test_nms_golden_bit_exact(case):
    from yolov3_b200.nms import non_max_suppression

    g = np.load(G / "nms_cases.npz")
    kw = ast.literal_eval(str(g[f"{case}/kw"]))
    pred = torch.from_numpy(g[str(g[f"{case}/pred_key"])]).cuda()
    outs, srcs = non_max_suppression(pred, return_src=True, **kw)
    for xi, (o, s) in enumerate(zip(outs, srcs)):
        ref = g[f"{case}/out{xi}"]
        assert o.shape == ref.shape, (case, xi, o.shape, ref.shape)
        assert np.array_equal(o.cpu().numpy(), ref), (case, xi)
        assert np.array_equal(s.cpu().numpy().astype(np.int64), g[f"{case}/src{xi}"])


@pytest.mark.parametrize("conf,iou,ml", [(0.25, 0.45, False), (0.001, 0.6, False), (0.05, 0.45, True), (0.001, 0.6, True)])
def test_nms_full_size_vs_oracle(conf, iou, ml):
    from yolov3_b200.nms import non_max_suppression

    pred = O.synth_predictions(2, n_rows=25200, nc=80, seed=3)
    outs, srcs = non_max_suppression(pred.cuda(), conf, iou, multi_label=ml, max_det=300, return_src=True)
    ref, rsrc = O.non_max_suppression(pred, conf, iou, multi_label=ml, max_det=300)
    for o, s, r, rs in zip(outs, srcs, ref, rsrc):
        assert np.array_equal(o.cpu().numpy(), r)
        assert np.array_equal(s.cpu().numpy().astype(np.int64), rs)


@pytest.mark.parametrize("conf,iou,ml", [(0.25, 0.45, False), (0.001, 0.6, False)])
def test_nms_benchmarked_batch_of_32(conf, iou, ml):
    """BASELINE config 5 at its full size (32 x 25200 x 85, the batch bench.py times): every image of the batch gets exactly
    the rows it gets as the only image of a call (the per-image segments of the batched pipeline do not leak into each other),
    and the first and the last image match the oracle bit for bit."""
    from yolov3_b200.nms import non_max_suppression

    pred = O.synth_predictions(32, n_rows=25200, nc=80, seed=3)
    dev = pred.cuda()
    outs = non_max_suppression(dev, conf, iou, multi_label=ml, max_det=300)
    assert len(outs) == 32
    for i in (0, 13, 31):
        alone = non_max_suppression(dev[i:i + 1], conf, iou, multi_label=ml, max_det=300)[0]
        assert torch.equal(outs[i], alone), i
    ref, _ = O.non_max_suppression(pred[[0, 31]], conf, iou, multi_label=ml, max_det=300)
    assert np.array_equal(outs[0].cpu().numpy(), ref[0]) and np.array_equal(outs[31].cpu().numpy(), ref[1])


def test_nms_properties_and_errors():
    from yolov3_b200.nms import non_max_suppression

    pred = O.synth_predictions(4, n_rows=25200, nc=80, seed=21).cuda()
    outs = non_max_suppression(pred, 0.25, 0.45, max_det=1000)
    again = non_max_suppression(pred, 0.25, 0.45, max_det=1000)
    for o, a in zip(outs, again):
        assert torch.equal(o, a)                              # deterministic
        assert (o[:-1, 4] >= o[1:, 4]).all()                  # sorted by confidence
        assert (o[:, 4] > 0.25).all() and o.shape[0] <= 1000
    # idempotence: NMS of the kept boxes (as predictions with obj=1, one-hot class) keeps all of them
    o = outs[0]
    p2 = torch.zeros(1, o.shape[0], 85, device="cuda")
    p2[0, :, 0] = (o[:, 0] + o[:, 2]) / 2
    p2[0, :, 1] = (o[:, 1] + o[:, 3]) / 2
    p2[0, :, 2] = o[:, 2] - o[:, 0]
    p2[0, :, 3] = o[:, 3] - o[:, 1]
    p2[0, :, 4] = 1.0
    p2[0, torch.arange(o.shape[0]), 5 + o[:, 5].long()] = o[:, 4]
    o2 = non_max_suppression(p2, 0.25, 0.45, max_det=1000)[0]
    assert o2.shape[0] >= o.shape[0] - 2  # re-derived boxes differ by 1 ulp; allow borderline pairs
    with pytest.raises(AssertionError):
        non_max_suppression(pred, conf_thres=1.5)
    with pytest.raises(AssertionError):
        non_max_suppression(pred, iou_thres=-0.1)
    # tuple input (inference_out, loss_out) like val.py passes
    t = non_max_suppression((pred, None), 0.25, 0.45)
    assert all(torch.equal(a, b[:300]) for a, b in zip(t, outs))
    # empty result
    e = non_max_suppression(pred, 1.0, 0.45)
    assert all(x.shape == (0, 6) for x in e)
