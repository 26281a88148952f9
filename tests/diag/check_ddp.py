"""2-rank hardware check of the overlapped data-parallel exchange (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/diag/check_ddp.py
Each rank trains on its own shard.  (1) A backward under ``ddp.no_sync()`` gives the rank-local gradient buffers; their mean over
ranks (all_gather on the host side of the test) is what DistributedDataParallel must produce (utils/torch_utils.py:60-72,
train.py:405-406).  (2) The same step with the exchange enabled — four bucketed NCCL all-reduces launched from inside the
backward on a side stream, 1/world folded in afterwards — must leave exactly that in EVERY parameter's .grad on EVERY rank.
(3) Two fused optimizer steps (clip + SGD + EMA) keep the replicas identical, with and without CUDA graphs."""
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
    from yolov3_b200.optim import SGD
    from yolov3_b200.train import TrainEngine

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    cfg = ROOT / "yolov3_b200" / "cfg" / "yolov3.yaml"
    hyp = O.scaled_hyp()
    ok = True
    for graphs in (False, True):
        TrainEngine.use_graphs = graphs
        TrainEngine.deterministic = True
        m = Model(cfg)
        m.load_state_dict(O.init_params(cfg, seed=rank))  # ranks start DIFFERENT: DDP's constructor broadcast must fix that
        m.hyp = hyp
        m.train()
        ddp = parallel.DDP(m)
        st = m.store()
        chk = st.P.double().sum()
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(lo == hi)
        x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(10 + rank)).cuda()
        t = O.synth_targets(2, seed=20 + rank).cuda()
        loss_fn = ComputeLoss(m)

        def backward(sync):
            m.zero_grad()
            loss, _ = loss_fn(m(x), t)
            loss = parallel.scale_loss(loss)
            if sync:
                loss.backward()
            else:
                with ddp.no_sync():
                    loss.backward()
            torch.cuda.synchronize()

        for _ in range(3 if graphs else 1):  # graphs: eager warm-up, capture, replay
            backward(False)
        local = st.G.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = sum(g.double() for g in gathered) / world
        for _ in range(3 if graphs else 1):
            backward(True)
        assert ddp.pending_average
        ddp.finish()
        got = st.G.double()
        err = float((got - want).abs().max() / want.abs().max())
        same = got.sum().clone()
        lo, hi = same.clone(), same.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= err < 1e-6 and bool(lo == hi)
        if rank == 0:
            print(f"graphs={graphs}: exchanged gradient vs mean of local gradients: max rel err {err:.2e}; identical on all ranks: {bool(lo == hi)}; "
                  f"buckets {[(b - a) * 4 // 1000000 for a, b in next(iter(m._train_engines.values())).buckets]} MB")
        # two fused steps keep the replicas identical
        opt = SGD(m, lr=0.01, momentum=0.937, weight_decay=5e-4, nesterov=True, max_norm=10.0)
        for _ in range(2):
            loss, _ = loss_fn(m(x), t)
            parallel.scale_loss(loss).backward()
            opt.step()
            opt.zero_grad()
        torch.cuda.synchronize()
        chk = st.P[:st.n_train].double().sum()  # trainable range: BatchNorm running statistics stay rank-local (no --sync-bn)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(lo == hi) and bool(torch.isfinite(chk))
        if rank == 0:
            print(f"graphs={graphs}: parameters identical on all ranks after 2 fused steps: {bool(lo == hi)}")
        del m, opt, ddp
        torch.cuda.empty_cache()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DDP_OK" if int(flag) else "DDP_FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
