"""2-rank hardware check of --sync-bn (run under torchrun, one rank per GPU):
rank r runs a training step on HALF of a batch with parallel.convert_sync_batchnorm(); rank 0 also runs the full batch
without it.  The synchronised running statistics must equal the full-batch ones, and the all-reduced (averaged, loss x
world) gradients must match the full-batch gradients to the step-to-step reproducibility of the bf16 pipeline.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/diag/check_syncbn.py"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import yolo_oracle as O

    from yolov3_b200 import parallel
    from yolov3_b200.loss import ComputeLoss
    from yolov3_b200.model import Model

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    cfg = ROOT / "yolov3_b200" / "cfg" / "yolov3.yaml"
    params = O.init_params(cfg, seed=0)
    hyp = O.scaled_hyp()
    bs = 4 * world
    x = torch.rand(bs, 3, 128, 128, generator=torch.Generator().manual_seed(3))
    t = O.synth_targets(bs, seed=2)

    def step(images, targets, sync):
        m = Model(cfg)
        m.load_state_dict(params)
        m.hyp = hyp
        if sync:
            parallel.convert_sync_batchnorm(m)
        m.train()
        loss, _ = ComputeLoss(m)(m(images.cuda()), targets.cuda())
        if sync:
            loss = parallel.scale_loss(loss)
        loss.backward()
        P = m.device_params()
        ps = [P[k] for k in sorted(P) if P[k].grad is not None]
        if sync:
            parallel.allreduce_gradients(ps)
        torch.cuda.synchronize()
        return m, {k: P[k].grad.clone() for k in sorted(P) if P[k].grad is not None}, {k: v.detach().clone() for k, v in P.items() if "running" in k}

    lo = rank * 4
    sel = (t[:, 0] >= lo) & (t[:, 0] < lo + 4)
    tl = t[sel].clone()
    tl[:, 0] -= lo
    _, g_sync, rs_sync = step(x[lo:lo + 4], tl, True)
    ok = True
    if rank == 0:
        _, g_full, rs_full = step(x, t, False)
        worst_rs = max(float((rs_sync[k] - rs_full[k]).abs().max() / rs_full[k].abs().max().clamp_min(1e-6)) for k in rs_full)
        errs = sorted(float((g_sync[k] - g_full[k]).norm() / g_full[k].norm().clamp_min(1e-30)) for k in g_full)
        med = errs[len(errs) // 2]
        print(f"sync-bn vs full batch: running stats worst rel diff {worst_rs:.2e}; gradient rel-L2 median {med:.3f} max {errs[-1]:.3f}")
        ok = worst_rs < 2e-2 and med < 0.25
        print("SYNCBN_OK" if ok else "SYNCBN_FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
