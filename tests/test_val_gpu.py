"""Validation matching on the device (y3_val_match) against the goldens produced by the reference's own
val.process_batch (val.py:147-188; tests/golden/make_golden.py gen_val) — bit-exact boolean matrices — and, on larger random
cases and a whole padded batch, against the oracle restatement."""
from pathlib import Path

import numpy as np
import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"


def _cases():
    g = np.load(G / "val_cases.npz")
    return sorted({k.split("/")[0] for k in g.files if "/" in k})


@pytest.mark.parametrize("case", _cases())
def test_process_batch_golden_exact(case):
    from yolov3_b200.val import process_batch

    g = np.load(G / "val_cases.npz")
    det, lab = torch.from_numpy(g[f"{case}/det"]).cuda(), torch.from_numpy(g[f"{case}/lab"]).cuda()
    iouv = torch.from_numpy(g["iouv"]).cuda()
    got = process_batch(det, lab, iouv)
    assert got.dtype == torch.bool and got.shape == (det.shape[0], 10) and got.device == iouv.device
    assert np.array_equal(got.cpu().numpy(), g[f"{case}/correct"])


def test_process_batch_batched_equals_per_image_oracle():
    """A padded batch (nms_batched layout: [bs, max_det, 6] + counts, collated labels (image, cls, xyxy)) in ONE launch equals
    the reference's per-image loop (val.py:372-388); rows beyond an image's count are False; > 256 detections per image."""
    from yolov3_b200.val import process_batch_batched

    bs, max_det = 5, 400
    iouv = torch.linspace(0.5, 0.95, 10)
    det = torch.zeros(bs, max_det, 6)
    counts = torch.tensor([400, 0, 37, 300, 1], dtype=torch.int32)
    labs, refs = [], []
    for i in range(bs):
        d, l = O.synth_val_case(int(counts[i]), [30, 4, 0, 55, 9][i], 5, seed=50 + i, jitter=6.0)
        det[i, : counts[i]] = d
        labs.append(torch.cat((torch.full((l.shape[0], 1), float(i)), l), 1))
        refs.append(O.process_batch(d, l, iouv) if l.shape[0] and d.shape[0] else torch.zeros(d.shape[0], 10, dtype=torch.bool))
    perm = torch.randperm(sum(x.shape[0] for x in labs), generator=torch.Generator().manual_seed(1))
    labels = torch.cat(labs, 0)
    # the image column, not the row order across images, selects an image's labels — but order INSIDE an image is the tie rule,
    # so shuffle whole images' blocks only
    labels = torch.cat([labs[j] for j in (3, 0, 4, 1, 2)], 0)
    del perm
    got = process_batch_batched(det.cuda(), counts.cuda(), labels.cuda(), iouv.cuda())
    assert got.shape == (bs, max_det, 10)
    for i in range(bs):
        n = int(counts[i])
        assert torch.equal(got[i, :n].cpu(), refs[i]), i
        assert not got[i, n:].any()


def test_process_batch_edges():
    from yolov3_b200.val import process_batch

    iouv = torch.linspace(0.5, 0.95, 10).cuda()
    det, lab = O.synth_val_case(10, 3, 2, seed=9)
    assert process_batch(det[:0].cuda(), lab.cuda(), iouv).shape == (0, 10)
    assert not process_batch(det.cuda(), lab[:0].cuda(), iouv).any()
    # duplicate labels (bit-equal IoU for one detection): the lower label index is the match, exactly one detection wins it
    lab2 = torch.cat((lab[:1], lab[:1]), 0)
    d2 = torch.cat((lab[:1, 1:], torch.tensor([[0.9, lab[0, 0]]])), 1).repeat(3, 1)
    out = process_batch(d2.cuda(), lab2.cuda(), iouv)
    assert out[0].all() and not out[1:].any()
    with pytest.raises(ValueError):
        process_batch(det.cuda(), torch.zeros(1025, 5).cuda(), iouv)
