"""Device letterbox / pre-processing (y3_letterbox_u8) against the oracle restatement of the reference's letterbox
(utils/augmentations.py:104-134 = cv2.resize INTER_LINEAR + constant border; pinned against cv2 and the reference itself in
tests/test_oracle_golden.py) — bit-exact uint8 images, identical ratio / padding — and the LoadImages layout step
(utils/dataloaders.py:308-310)."""
import numpy as np
import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu


def _img(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base[..., 0] = (base[..., 0] // 2 + (xx * 255 // max(w - 1, 1)) // 2).astype(np.uint8)  # gradients + noise
    base[..., 2] = (base[..., 2] // 2 + (yy * 255 // max(h - 1, 1)) // 2).astype(np.uint8)
    return base


@pytest.mark.parametrize("hw", [(1080, 810), (720, 1280), (375, 500), (480, 640), (100, 133), (1280, 960), (640, 640), (333, 1000), (17, 23)])
@pytest.mark.parametrize("kw", [dict(auto=True), dict(auto=False), dict(auto=False, scaleFill=True), dict(auto=True, scaleup=False),
                                dict(new_shape=(384, 640), auto=False), dict(new_shape=320, auto=True, stride=64)])
def test_letterbox_bit_exact(hw, kw):
    from yolov3_b200.preprocess import letterbox

    im = _img(*hw, seed=hw[0] * 7 + hw[1])
    ref, r_ratio, r_pad = O.letterbox(im.copy(), **kw)
    got, ratio, pad = letterbox(torch.from_numpy(im).cuda(), **kw)
    assert got.shape == ref.shape and ratio == r_ratio and tuple(pad) == tuple(r_pad)
    assert np.array_equal(got.cpu().numpy(), ref)


def test_preprocess_layout_and_model_input():
    """preprocess() == letterbox -> transpose((2,0,1))[::-1] (BGR->RGB, CHW), written into an engine's uint8 input batch; the
    model's forward on that batch equals its forward on the host-prepared batch."""
    from pathlib import Path

    from yolov3_b200.model import Model
    from yolov3_b200.preprocess import preprocess

    ims = [_img(300, 400, 1), _img(300, 400, 2)]
    cfg = Path(__file__).resolve().parents[1] / "yolov3_b200" / "cfg" / "yolov3-tiny.yaml"
    m = Model(cfg)
    m.load_state_dict(O.init_params(cfg, seed=0))
    ref = np.stack([O.preprocess(im, 416, stride=32, auto=True) for im in ims])
    batch = torch.empty(2, 3, ref.shape[2], ref.shape[3], dtype=torch.uint8, device="cuda")
    for i, im in enumerate(ims):
        out, ratio, pad = preprocess(torch.from_numpy(im).cuda(), 416, stride=32, auto=True, out=batch[i])
        assert out.data_ptr() == batch[i].data_ptr()
    assert np.array_equal(batch.cpu().numpy(), ref)
    z_dev, _ = m(batch)
    z_host, _ = m(torch.from_numpy(ref).cuda())
    assert torch.equal(z_dev, z_host)
    with pytest.raises(AssertionError):
        preprocess(torch.from_numpy(ims[0]).cuda(), 416, out=torch.empty(3, 10, 10, dtype=torch.uint8, device="cuda"))
