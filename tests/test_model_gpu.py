"""Whole-graph parity of yolov3_b200.Model against the CPU oracle and the committed reference goldens.

Stated tolerance (bf16 storage of activations/weights, fp32 accumulation): relative L2 error
  <= 2e-2 on raw logits p_i and on decoded z vs the fp32 oracle (== the reference forward, tests/golden/forward_*.npz),
  <= 4e-3 vs the oracle that emulates the same bf16 storage rounding (isolates kernel errors from format error)."""
from pathlib import Path

import numpy as np
import pytest
import torch

import yolo_oracle as O

pytestmark = pytest.mark.gpu
CFG = Path(__file__).resolve().parents[1] / "yolov3_b200" / "cfg"
G = Path(__file__).parent / "golden"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("name,shape", [("yolov3-tiny", (2, 3, 96, 128)), ("yolov3", (2, 3, 64, 96)),
                                        ("yolov3-spp", (1, 3, 64, 64)), ("yolov3", (1, 3, 160, 128))])
def test_forward_vs_oracle(name, shape):
    from yolov3_b200.model import Model

    params = O.init_params(CFG / f"{name}.yaml", seed=0)
    m = Model(CFG / f"{name}.yaml")
    m.load_state_dict(params)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(11))
    z, raw = m(x.cuda())
    torch.cuda.synchronize()
    e = m.engine(shape[0], shape[2], shape[3])
    e.check_errors()
    o32 = O.OracleModel(CFG / f"{name}.yaml", params=params, fused=True)
    o16 = O.OracleModel(CFG / f"{name}.yaml", params=params, fused=True, act_dtype=torch.bfloat16,
                        weight_dtype=torch.bfloat16)
    with torch.no_grad():
        z32, raw32 = o32(x)
        z16, raw16 = o16(x)
    assert z.shape == z32.shape and all(a.shape == b.shape for a, b in zip(raw, raw32))
    for a, b in zip(raw, raw16):
        assert rel_l2(a, b) <= 4e-3, ("bf16-emulating oracle", rel_l2(a, b))
    for a, b in zip(raw, raw32):
        assert rel_l2(a, b) <= 2e-2, ("fp32 oracle", rel_l2(a, b))
    assert rel_l2(z, z16) <= 4e-3 and rel_l2(z, z32) <= 2e-2


@pytest.mark.parametrize("name", ["yolov3-tiny", "yolov3", "yolov3-spp"])
def test_forward_vs_reference_golden(name):
    from yolov3_b200.model import Model

    g = np.load(G / f"forward_{name}.npz")
    params = O.init_params(CFG / f"{name}.yaml", seed=int(g["param_seed"]))
    m = Model(CFG / f"{name}.yaml")
    m.load_state_dict(params)
    assert m.save == list(g["save"]) and np.array_equal(m.stride.numpy(), g["stride"])
    ci = 0
    while f"z{ci}" in g:
        bs, c, h, w = (int(v) for v in g[f"x{ci}_shape"])
        x = torch.rand(bs, c, h, w, generator=torch.Generator().manual_seed(int(g[f"x{ci}_seed"])))
        z, raw = m(x.cuda())
        assert rel_l2(z, torch.from_numpy(g[f"z{ci}"])) <= 2e-2
        for li, r in enumerate(raw):
            assert rel_l2(r, torch.from_numpy(g[f"raw{ci}_{li}"])) <= 2e-2
        ci += 1


def test_uint8_input_and_graph_replay():
    from yolov3_b200.model import Model

    m = Model(CFG / "yolov3-tiny.yaml")
    m.load_state_dict(O.init_params(CFG / "yolov3-tiny.yaml", seed=0))
    xi = torch.randint(0, 256, (2, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    z8, _ = m(xi.cuda())
    zf, _ = m((xi.float() / 255).cuda())
    assert rel_l2(z8, zf) < 1e-6
    e = m.engine(2, 64, 64, torch.float32)
    e.static_in.copy_((xi.float() / 255).cuda())
    e.capture()
    zg, _ = e.replay()
    torch.cuda.synchronize()
    assert torch.equal(zg, zf)


def test_decode_matches_oracle():
    from yolov3_b200.detect import decode

    g = torch.Generator().manual_seed(8)
    anchors = torch.tensor([[[1.25, 1.625], [2.0, 3.75], [4.125, 2.875]], [[1.875, 3.8125], [3.875, 2.8125], [3.6875, 7.4375]]])
    stride = torch.tensor([8.0, 16.0])
    raw = [torch.randn(2, 3, 6, 10, 85, generator=g) * 3, torch.randn(2, 3, 3, 5, 85, generator=g) * 3]
    z = decode([r.cuda() for r in raw], anchors, stride)
    ref = O.decode(raw, anchors, stride)
    assert torch.allclose(z.cpu(), ref, rtol=2e-6, atol=1e-6)


def test_no_cpu_path():
    from yolov3_b200.model import Model

    m = Model(CFG / "yolov3-tiny.yaml")
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))

