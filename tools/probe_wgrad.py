"""Probe of the tcgen05 wgrad kernel's MN-major descriptor strides: runs y3_conv_wgrad for a few shapes under each
Y3_WGRAD_VARIANT (own subprocess: a wrong descriptor may trap) and prints the rel-L2 error against torch autograd.
  python tools/probe_wgrad.py            -> variants 0,1,2 + the warp-MMA kernel as control
  python tools/probe_wgrad.py --one      -> worker (uses the environment as is)"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CASES = [(64, 128, 3, 2, 12, 20), (128, 64, 1, 2, 16, 16), (32, 64, 3, 1, 16, 16), (256, 256, 3, 2, 20, 20), (384, 128, 1, 1, 20, 20),
         (64, 32, 1, 1, 24, 24)]


def worker():
    import torch
    import torch.nn.functional as F

    from yolov3_b200 import train_ops as T
    from yolov3_b200.tensors import PaddedNHWC

    for ci, co, k, n, h, w in CASES:
        g = torch.Generator().manual_seed(ci + co)
        x = torch.randn(n, ci, h, w, generator=g).bfloat16().float()
        dy = torch.randn(n, co, h, w, generator=g).bfloat16().float()
        wt = torch.zeros(co, ci, k, k, requires_grad=True)
        F.conv2d(x, wt, padding=k // 2).backward(dy)
        xp = PaddedNHWC.zeros(n, h, w, ci).load_nchw(x.cuda())
        dyp = PaddedNHWC.zeros(n, h, w, co).load_nchw(dy.cuda())
        dw = torch.zeros(co, ci, k, k, device="cuda")
        T.conv_wgrad(dyp, xp, dw, k)
        torch.cuda.synchronize()
        err = float((dw.cpu() - wt.grad).norm() / wt.grad.norm())
        print(f"  ci={ci} co={co} k={k} n={n} {h}x{w}: rel-L2 {err:.3e}", flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        worker()
    else:
        for name, env in (("warp-mma control", {"Y3_WGRAD_TC": "0"}), ("tc variant 0", {"Y3_WGRAD_VARIANT": "0"}),
                          ("tc variant 1", {"Y3_WGRAD_VARIANT": "1"}), ("tc variant 2", {"Y3_WGRAD_VARIANT": "2"})):
            print(name, flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, "--one"], env={**os.environ, **env}, capture_output=True, text=True, timeout=120)
                print(r.stdout, end="")
                if r.returncode:
                    print("  exit", r.returncode, r.stderr.strip().splitlines()[-1][:200] if r.stderr.strip() else "")
            except subprocess.TimeoutExpired:
                print("  TIMEOUT")
