"""Host-side checks of kernel index arithmetic: the headers the CUDA kernels compile are also compiled with g++ into a small
test library and driven with the kernel's loop structure (tests/emul/*.cpp), so the element -> address mapping of those
kernels is pinned against the oracle without a GPU.  (The GPU tests compare the real kernels' results.)"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import yolo_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "libemul.so"
    srcs = sorted(str(p) for p in (ROOT / "tests" / "emul").glob("*.cpp"))
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out), *srcs], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(out))


@pytest.mark.parametrize("bs,na,nc,shapes,grid", [
    (2, 3, 80, [(8, 8), (4, 4), (16, 2)], 7),      # yolov3 head: no 85, head_ld 256
    (3, 3, 80, [(20, 20), (12, 4)], 1000),         # grid > tiles
    (1, 3, 1, [(4, 8), (4, 4)], 3),                # no 6 -> 18 channels in head_ld 32
    (2, 2, 20, [(8, 6)], 5),                       # no 25, na 2 -> 50 channels in head_ld 64
    (1, 5, 46, [(4, 4)], 2),                       # 255 of 256 channels with another (na, no) split
])
def test_staged_decode_index_arithmetic(emul, bs, na, nc, shapes, grid):
    no = nc + 5
    ld = 32
    while ld < na * no:
        ld *= 2
    g = torch.Generator().manual_seed(bs * 100 + na * 10 + nc)
    nl = len(shapes)
    heads, raws = [], []
    for ny, nx in shapes:
        h = torch.randn(bs * ny * nx, ld, generator=g) * 2.0
        heads.append(h.contiguous())
        # the reference-layout view of the head buffer (model.py: raw = head.view(n,h,w,ld)[..., :na*no].unflatten.permute)
        raws.append(h.view(bs, ny, nx, ld)[..., : na * no].unflatten(-1, (na, no)).permute(0, 3, 1, 2, 4).contiguous())
    stride = [8.0 * 2 ** i for i in range(nl)]
    anchors_px = torch.rand(nl, na, 2, generator=g) * 100 + 4
    anchors_grid = anchors_px / torch.tensor(stride).view(nl, 1, 1)
    z_ref = O.decode(raws, anchors_grid, torch.tensor(stride)).numpy()
    rows = sum(na * ny * nx for ny, nx in shapes)
    z = np.full((bs, rows, no), np.nan, np.float32)
    fn = emul.decode2_emul
    fn.restype = C.c_int
    ny_a = (C.c_int * nl)(*[s[0] for s in shapes])
    nx_a = (C.c_int * nl)(*[s[1] for s in shapes])
    st_a = (C.c_float * nl)(*stride)
    anc = np.ascontiguousarray((anchors_grid * torch.tensor(stride).view(nl, 1, 1)).numpy().astype(np.float32))
    hp = (C.c_void_p * nl)(*[h.data_ptr() for h in heads])
    rc = fn(nl, bs, na, no, ld, ny_a, nx_a, st_a, anc.ctypes.data_as(C.c_void_p), hp, z.ctypes.data_as(C.c_void_p), grid)
    assert rc == 0, f"emulation self-check {rc}"
    assert np.isfinite(z).all()
    np.testing.assert_allclose(z, z_ref, rtol=2e-6, atol=2e-6)
