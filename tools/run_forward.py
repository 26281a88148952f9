"""Target for ncu: a few eager (non-graph) yolov3 bs32 640 forwards.  Usage: python tools/run_forward.py [iters] [bs]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import BS, IMG, build_model  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bs = int(sys.argv[2]) if len(sys.argv) > 2 else BS
dev = torch.device("cuda", 0)
m = build_model(dev)
e = m.engine(bs, IMG, IMG, torch.float32)
x = torch.rand(bs, 3, IMG, IMG, device=dev)
for _ in range(iters):
    e.run(x)
torch.cuda.synchronize()
e.check_errors()
print("done", e.n_ops, "launches per forward")
