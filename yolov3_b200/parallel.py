"""Data-parallel gradient exchange — the ONE collective on the path.

The reference wraps the model in ``DistributedDataParallel`` (utils/torch_utils.py:60-72, train.py:323): one process
per GPU, each rank runs forward/backward on its shard, gradients are averaged with a bucketed NCCL all-reduce, and
``loss *= WORLD_SIZE`` (train.py:405-406) undoes the averaging because the loss is already scaled by the rank batch size.
Here the backward is a single autograd node (``TrainFn``), so there is nothing to overlap bucket-by-bucket with; the
exchange is one flat all-reduce of the 62 M fp32 gradients (248 MB) over NCCL / NVLink, issued right after ``backward()``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def scale_loss(loss: torch.Tensor) -> torch.Tensor:
    """train.py:405-406: ``loss *= WORLD_SIZE`` (gradients are averaged between ranks afterwards)."""
    return loss * world_size()


def allreduce_gradients(params, average: bool = True, flat: torch.Tensor | None = None) -> torch.Tensor | None:
    """All-reduce (mean) the ``.grad`` of ``params`` across ranks in ONE flat buffer.  Returns the flat buffer so the
    caller can keep it alive / reuse it.  No-op for a single process."""
    ps = [p for p in params if p.grad is not None]
    w = world_size()
    if w == 1 or not ps:
        return flat
    n = sum(p.grad.numel() for p in ps)
    if flat is None or flat.numel() != n or flat.device != ps[0].grad.device:
        flat = torch.empty(n, dtype=torch.float32, device=ps[0].grad.device)
    off = 0
    for p in ps:
        k = p.grad.numel()
        flat[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(w)
    off = 0
    for p in ps:
        k = p.grad.numel()
        p.grad.copy_(flat[off:off + k].view_as(p.grad))
        off += k
    return flat


def broadcast_parameters(params, src: int = 0):
    """DDP's constructor broadcast: every rank starts from rank ``src``'s parameters (train.py:323)."""
    if world_size() == 1:
        return
    for p in params:
        dist.broadcast(p.data, src)


def convert_sync_batchnorm(model):
    """``torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)`` of the reference's ``--sync-bn`` switch (train.py:270-272):
    the training engine all-reduces every BatchNorm's (sum, sum of squares) in the forward and (sum dz, sum dz*xhat) in the
    backward, so batch statistics cover all ranks' images.  gamma/beta gradients stay rank-local sums (the gradient
    all-reduce averages them like every other parameter), exactly as torch's SyncBatchNorm does under DDP."""
    model.sync_bn = True
    model._train_engines.clear()
    return model
