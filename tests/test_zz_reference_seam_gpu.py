"""The seam, exercised with the REFERENCE's own scripts: the staged, unmodified reference (baseline/_ref,
oracle/stage_reference.py) is imported through the shim and its ``detect.run`` is executed with ``DetectMultiBackend``,
``non_max_suppression`` and ``scale_boxes`` swapped for ours; the val.py loop body and the train.py optimizer/EMA objects run
against the nn.Module facade.  Skipped when the staged copy is absent (fresh clone without /root/reference)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
CFG = ROOT / "yolov3_b200" / "cfg"
sys.path.insert(0, str(ROOT / "oracle"))
import ref_shim  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not staged (baseline/_ref)")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _confident_params(cfg, seed=0):
    """Random-init weights whose objectness / class biases are raised so that detections exist at ordinary thresholds (the
    shipped init gives conf ~ 3e-5 everywhere: nothing to compare)."""
    p = O.init_params(cfg, seed=seed)
    for k in p:
        if ".m." in k and k.endswith(".bias"):
            b = p[k].view(3, -1)
            b[:, 4] += 7.0   # sigmoid(-3.9 .. -5.3 + 7) = 0.85 .. 0.96
            b[:, 5:] += 5.0  # sigmoid(-4.9 + 5 +- noise) ~ 0.5: obj * cls crosses 0.25 for a share of the classes
    return p


@needs_ref
def test_reference_detect_py_runs_on_our_backend(tmp_path):
    """BASELINE config 1 plumbing with the backend swapped in: the reference's detect.run (detect.py:47-235) on data/images with
    our DetectMultiBackend / non_max_suppression / scale_boxes; then, image by image, our forward vs the reference model's
    (rel-L2 <= 2e-2) and — on the reference's own predictions — our NMS + scale_boxes vs the reference's (bit-exact)."""
    ref_shim.install()
    import detect as ref_detect  # reference script
    from models.yolo import Model as RefModel
    from utils.dataloaders import LoadImages
    from utils.general import non_max_suppression as ref_nms
    from utils.general import scale_boxes as ref_scale

    from yolov3_b200 import backend, boxes, nms
    from yolov3_b200.model import Model

    cfg = CFG / "yolov3-tiny.yaml"
    params = _confident_params(cfg)
    m = Model(cfg)
    m.load_state_dict(params)
    ckpt = tmp_path / "tiny_b200.pt"
    backend.save_checkpoint(m, ckpt)
    src = ref_shim.REFERENCE_ROOT / "data" / "images"
    saved = (ref_detect.DetectMultiBackend, ref_detect.non_max_suppression, ref_detect.scale_boxes)
    ref_detect.DetectMultiBackend, ref_detect.non_max_suppression, ref_detect.scale_boxes = \
        backend.DetectMultiBackend, nms.non_max_suppression, boxes.scale_boxes
    try:
        ref_detect.run(weights=str(ckpt), source=str(src), imgsz=(640, 640), conf_thres=0.25, iou_thres=0.45, max_det=50, device="0",
                       save_txt=True, save_conf=True, nosave=True, project=str(tmp_path), name="exp", exist_ok=True)
    finally:
        ref_detect.DetectMultiBackend, ref_detect.non_max_suppression, ref_detect.scale_boxes = saved
    labels = sorted((tmp_path / "exp" / "labels").glob("*.txt"))
    assert len(labels) == 2 and all(len(p.read_text().splitlines()) >= 1 for p in labels), labels
    # ---- parity of the swapped pieces on the same images
    rm = RefModel(str(ref_shim.REFERENCE_ROOT / "models" / "yolov3-tiny.yaml"))
    rm.load_state_dict(params, strict=False)
    rm = rm.eval()
    ours = backend.DetectMultiBackend(str(ckpt), device=torch.device("cuda"))
    for path, im, im0s, _, _ in LoadImages(str(src), img_size=(640, 640), stride=32, auto=True):
        x = torch.from_numpy(im).float()[None] / 255
        with torch.no_grad():
            z_ref = rm(x)[0]
        z = ours(x.cuda())[0]
        assert rel_l2(z, z_ref) <= 2e-2, path
        det_ref = ref_nms(z_ref.clone(), 0.25, 0.45, max_det=50)[0]
        det = nms.non_max_suppression(z_ref.cuda(), 0.25, 0.45, max_det=50)[0]
        # real-image predictions with saturating sigmoids contain bit-equal confidences (birthday collisions among ~20 k
        # candidates): the reference orders such ties by an unstable argsort, we by candidate index — compare row SETS
        # (rows sorted lexicographically), which is what "identical detections" means when the order is undefined
        a, b = det_ref.numpy(), det.cpu().numpy()
        assert a.shape == b.shape, (path, a.shape, b.shape)
        ka = np.lexsort(a.T[::-1])
        kb = np.lexsort(b.T[::-1])
        assert np.array_equal(a[ka], b[kb]), (path, np.abs(a[ka] - b[kb]).max())
        det = det_ref.cuda()  # continue with identical rows in identical order
        sa = ref_scale(x.shape[2:], det_ref[:, :4].clone(), im0s.shape)
        sb = boxes.scale_boxes(x.shape[2:], det[:, :4].clone(), im0s.shape)
        assert np.array_equal(sb.cpu().numpy(), sa.numpy())


@needs_ref
def test_val_loop_body_pieces_match_reference(tmp_path):
    """val.py:355-390 with our pieces on the reference's predictions: NMS (multi-label, conf 0.001, iou 0.6), scale_boxes to
    native space and process_batch give exactly the reference's ``correct`` matrix."""
    ref_shim.install()
    import val as ref_val
    from utils.general import non_max_suppression as ref_nms
    from utils.general import scale_boxes as ref_scale
    from utils.general import xywh2xyxy

    from yolov3_b200 import boxes, nms
    from yolov3_b200.val import process_batch

    pred = O.synth_predictions(2, n_rows=3000, nc=80, seed=5)
    targets = O.synth_targets(2, seed=4)
    h = w = 640
    shape0, ratio_pad = (480, 600), ((1.0667, 1.0667), (0.0, 64.0))
    targets_px = targets.clone()
    targets_px[:, 2:] *= torch.tensor((w, h, w, h))
    iouv = torch.linspace(0.5, 0.95, 10)
    import utils.general as G

    real_time = G.time.time
    G.time.time = lambda: 0.0  # the reference's wall-clock break (utils/general.py:675,746-748) would drop slow images
    try:
        ref_out = ref_nms(pred.clone(), 0.001, 0.6, multi_label=True, max_det=300)
    finally:
        G.time.time = real_time
    our_out = nms.non_max_suppression(pred.cuda(), 0.001, 0.6, multi_label=True, max_det=300)
    for si in range(2):
        labels = targets_px[targets_px[:, 0] == si, 1:]
        predn_ref = ref_out[si].clone()
        ref_scale((h, w), predn_ref[:, :4], shape0, ratio_pad)
        tbox = xywh2xyxy(labels[:, 1:5])
        ref_scale((h, w), tbox, shape0, ratio_pad)
        labelsn = torch.cat((labels[:, 0:1], tbox), 1)
        correct_ref = ref_val.process_batch(predn_ref, labelsn, iouv)
        predn = our_out[si].clone()
        assert np.array_equal(predn.cpu().numpy(), ref_out[si].numpy())
        boxes.scale_boxes((h, w), predn[:, :4], shape0, ratio_pad)
        tb = xywh2xyxy(labels[:, 1:5]).cuda()
        boxes.scale_boxes((h, w), tb, shape0, ratio_pad)
        correct = process_batch(predn, torch.cat((labels[:, 0:1].cuda(), tb), 1), iouv.cuda())
        assert np.array_equal(predn.cpu().numpy(), predn_ref.numpy())
        assert np.array_equal(correct.cpu().numpy(), correct_ref.numpy()), si


def _train_two_steps(model, opt_step, zero_grad, x, targets, loss_fn):
    losses = []
    for _ in range(2):
        pred = model(x)
        loss, _ = loss_fn(pred, targets)
        loss.backward()
        opt_step()
        zero_grad()
        losses.append(float(loss.detach()))
    return losses


@needs_ref
def test_facade_with_reference_optimizer_and_ema():
    """train.py's objects on the facade: ``smart_optimizer`` (utils/torch_utils.py:207-237) sorts our parameters into its three
    groups, ``ModelEMA`` (train.py:252) deep-copies and updates, ``clip_grad_norm_`` + ``optimizer.step()`` train the masters —
    and two steps land where our fused step (optim.SGD + ModelEMA) lands on an identically initialised model."""
    ref_shim.install()
    from ultralytics.utils.torch_utils import ModelEMA as RefEMA  # the shim's restatement of the third-party class
    from utils.torch_utils import smart_optimizer

    from yolov3_b200.loss import ComputeLoss
    from yolov3_b200.module import DetectionModel
    from yolov3_b200.optim import SGD, ModelEMA
    from yolov3_b200.train import TrainEngine

    TrainEngine.deterministic = True
    try:
        cfg = CFG / "yolov3.yaml"
        params = O.init_params(cfg, seed=0)
        x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)).cuda()
        targets = O.synth_targets(2, seed=2).cuda()
        hyp = O.scaled_hyp()
        # ---- A: facade + reference optimizer + reference-style EMA
        ma = DetectionModel(cfg)
        ma.load_state_dict(params, strict=False)
        ma.hyp = hyp
        names = {n for n, _ in ma.named_parameters()}
        assert "model.4.0.cv1.conv.weight" in names and "model.28.m.2.bias" in names and len(names) == 222
        opt = smart_optimizer(ma, "SGD", 0.01, 0.937, 5e-4)
        assert [len(g["params"]) for g in opt.param_groups] == [75, 75, 72]  # biases (72 BN + 3 head), decay weights, BN weights
        ema = RefEMA(ma)
        ma.train()
        loss_fn = ComputeLoss(ma)

        def step_a():
            torch.nn.utils.clip_grad_norm_(ma.parameters(), max_norm=10.0)
            opt.step()
            ema.update(ma)

        la = _train_two_steps(ma, step_a, opt.zero_grad, x, targets, loss_fn)
        # ---- B: plain Model + fused optimizer
        mb = DetectionModel(cfg)
        mb.load_state_dict(params, strict=False)
        mb.hyp = hyp
        mb.train()
        ema_b = ModelEMA(mb.core)
        opt_b = SGD(mb.core, lr=0.01, momentum=0.937, weight_decay=5e-4, nesterov=True, max_norm=10.0, ema=ema_b)
        lb = _train_two_steps(mb, opt_b.step, opt_b.zero_grad, x, targets, ComputeLoss(mb))
        assert la == lb, (la, lb)  # deterministic engines, identical updates after step 1 -> identical loss at step 2
        sa, sb = ma.state_dict(), mb.state_dict()
        for k in sa:
            if "num_batches_tracked" in k:
                continue
            assert torch.allclose(sa[k].float(), sb[k].float(), rtol=1e-5, atol=1e-7), k
        ea, eb = ema.ema.state_dict(), ema_b.state_dict()
        for k in eb:
            assert torch.allclose(ea[k].float().cpu(), eb[k], rtol=1e-5, atol=1e-7), k
        # eval-mode inference picks the trained weights up (in-place updates are detected through the store's version)
        ma.eval()
        mb.eval()
        za, zb = ma(x)[0], mb(x)[0]
        assert torch.equal(za, zb) and bool(torch.isfinite(za).all())
        # half(): fp16 outputs, masters rounded to fp16-representable values (train.py:317)
        zh = ma.half()(x.half())[0]
        assert zh.dtype == torch.float16
        w = ma.state_dict()["model.5.conv.weight"]
        assert torch.equal(w, w.half().float())
    finally:
        TrainEngine.deterministic = False
