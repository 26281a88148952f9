"""Parity of the BENCHMARKED graph instances: the 640x640 engines bench.py times — yolov3.yaml
bs 32 (BASELINE config 2) and yolov3-spp.yaml bs 8 (config 3, per-GPU shard) — against the CPU oracle (== the reference
forward, models/yolo.py:135-147), including the CUDA-graph replay bench.py uses.  At 640^2 every CTA-pair tile is full,
the deep layers run several waves and the SPP pools see a 20x20 map (models/common.py:281-290) — none of which the 64..160
pixel goldens exercise.

Stated tolerance (same as test_model_gpu.py): rel-L2 <= 2e-2 on raw maps / z vs the fp32 oracle, <= 4e-3 vs the oracle that
emulates bf16 storage.  The oracle runs only on the images compared (first and last of the batch: ~1 s of CPU each)."""
from pathlib import Path

import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu
CFG = Path(__file__).resolve().parents[1] / "yolov3_b200" / "cfg"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("name,bs", [("yolov3", 32), ("yolov3-spp", 8)])
def test_forward_640_benchmarked_instance(name, bs):
    from yolov3_b200.model import Model

    cfg = CFG / f"{name}.yaml"
    params = O.init_params(cfg, seed=0)
    m = Model(cfg)
    m.load_state_dict(params)
    x = torch.rand(bs, 3, 640, 640, generator=torch.Generator().manual_seed(21))
    xd = x.cuda()
    e = m.engine(bs, 640, 640, torch.float32)
    z, raw = e.run(xd)
    torch.cuda.synchronize()
    e.check_errors()
    assert z.shape == (bs, 25200, 85) and bool(torch.isfinite(z).all())
    z_eager = z.clone()
    raw_eager = [r.contiguous().clone() for r in raw]
    # the CUDA-graph replay bench.py times must reproduce the eager launch sequence bit for bit
    e.capture(xd)
    e.z.zero_()
    zg, _ = e.replay()
    torch.cuda.synchronize()
    e.check_errors()
    assert torch.equal(zg, z_eager)

    pick = [0, bs - 1]
    xs = x[pick]
    o32 = O.OracleModel(cfg, params=params, fused=True)
    o16 = O.OracleModel(cfg, params=params, fused=True, act_dtype=torch.bfloat16, weight_dtype=torch.bfloat16)
    with torch.no_grad():
        z32, raw32 = o32(xs)
        z16, raw16 = o16(xs)
    for li, (a, b32, b16) in enumerate(zip(raw_eager, raw32, raw16)):
        assert a.shape[1:] == b32.shape[1:]
        for j, i in enumerate(pick):  # per image: a wrong tile in one image must not hide in the batch norm
            assert rel_l2(a[i], b16[j]) <= 4e-3, (name, "raw vs bf16-emulating oracle", li, i, rel_l2(a[i], b16[j]))
            assert rel_l2(a[i], b32[j]) <= 2e-2, (name, "raw vs fp32 oracle", li, i, rel_l2(a[i], b32[j]))
    for j, i in enumerate(pick):
        assert rel_l2(z_eager[i], z16[j]) <= 4e-3 and rel_l2(z_eager[i], z32[j]) <= 2e-2
    # every image of the batch went through the same lowering: identical inputs give identical rows
    x2 = x.clone()
    x2[bs // 2] = x[0]
    z2, _ = e.run(x2.cuda())
    torch.cuda.synchronize()
    assert torch.equal(z2[bs // 2], z_eager[0])


@pytest.mark.parametrize("hw", [(20, 20), (13, 13), (14, 14), (5, 7), (40, 24)])
def test_spp_pool_cascade_vs_max_pool2d(hw):
    """The inference SPP block (models/common.py:281-290) is lowered as a 5x5 stride-1 cascade (5, 5o5 = 9, 5o5o5 = 13)
    into channel slices of one concat buffer (model.py: Engine._lower).  Exact vs F.max_pool2d 5/9/13 with -inf padding, on
    maps at and above the 13x13 window (the 20x20 P5 map of config 3 included) and below it (border-dominated)."""
    import torch.nn.functional as F

    from yolov3_b200 import ops
    from yolov3_b200.tensors import PaddedNHWC

    h, w = hw
    c = 64
    x = torch.randn(2, c, h, w, generator=torch.Generator().manual_seed(h * 100 + w)).bfloat16().float()
    cat = PaddedNHWC.zeros(2, h, w, 4 * c)
    cat.slice(0, c).load_nchw(x.cuda())
    for q in range(3):
        ops.maxpool(cat.slice(q * c, c), cat.slice((q + 1) * c, c), 5, 1, -2, False)
    got = cat.to_nchw().cpu()
    ref = torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in (5, 9, 13)], 1)
    assert torch.equal(got, ref)
    halo = cat.buf.float().clone()
    halo[:, 1:-1, 1:-1] = 0
    assert (halo == 0).all()


def test_spp_block_vs_oracle_20x20():
    """cv1 -> pools -> concat -> cv2 of the yolov3-spp SPP node on a 20x20 map through the whole-graph engine: the layer-12
    tap of the 640^2 spp engine is covered by test_forward_640_benchmarked_instance; this is the same block alone at bs 2
    against torch ops on identical bf16 operands."""
    import torch.nn.functional as F

    from yolov3_b200 import ops
    from yolov3_b200.tensors import PaddedNHWC

    g = torch.Generator().manual_seed(77)
    n, c1, c2, h, w = 2, 128, 64, 20, 20
    c_ = c1 // 2
    x = torch.randn(n, c1, h, w, generator=g).bfloat16().float()
    w1 = (torch.randn(c_, c1, 1, 1, generator=g) / c1 ** 0.5).bfloat16().float()
    b1 = torch.randn(c_, generator=g) * 0.1
    w2 = (torch.randn(c2, 4 * c_, 1, 1, generator=g) / (4 * c_) ** 0.5).bfloat16().float()
    b2 = torch.randn(c2, generator=g) * 0.1
    xin = PaddedNHWC.zeros(n, h, w, c1).load_nchw(x.cuda())
    cat = PaddedNHWC.zeros(n, h, w, 4 * c_)
    ops.conv_bn_act(xin, *ops.pack_conv_weight(w1, b1), c_, 1, 1, ops.ACT_SILU, out=cat.slice(0, c_))
    for q in range(3):
        ops.maxpool(cat.slice(q * c_, c_), cat.slice((q + 1) * c_, c_), 5, 1, -2, False)
    y = ops.conv_bn_act(cat, *ops.pack_conv_weight(w2, b2), c2, 1, 1, ops.ACT_SILU)
    t = F.silu(F.conv2d(x, w1, b1)).bfloat16().float()
    t = torch.cat([t] + [F.max_pool2d(t, k, 1, k // 2) for k in (5, 9, 13)], 1)
    ref = F.silu(F.conv2d(t, w2, b2))
    got = y.to_nchw().cpu()
    assert torch.allclose(got, ref, atol=2e-2, rtol=1e-2), (got - ref).abs().max()
