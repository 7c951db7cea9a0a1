"""Distributed helpers — same names/behaviour as /root/reference/src/d_fine/dist_utils.py:13-205
for the train-step path (process-group init from torchrun's env, rank helpers, scalar
broadcast, dict reduce, barrier).  Eval-time object gathers are out of scope.

One process per GPU; NCCL over NVLink/NVSwitch carries exactly two things per step: the
gradient buckets (``wrap_ddp``) and one 2-float all-reduce of the loss normalisers.
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.distributed as dist


def is_dist_available_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def init_distributed_mode() -> None:
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        warnings.warn("DDP is enabled in config but RANK/WORLD_SIZE are not set; launch with torchrun")
        return
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        try:
            dist.init_process_group(backend="nccl", init_method="env://", device_id=torch.device("cuda", local_rank))
        except TypeError:
            dist.init_process_group(backend="nccl", init_method="env://")
    else:
        dist.init_process_group(backend="gloo", init_method="env://")


def cleanup_distributed() -> None:
    if is_dist_available_and_initialized():
        dist.destroy_process_group()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_available_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_available_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def get_local_rank() -> int:
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    if "RANK" in os.environ and torch.cuda.is_available():
        return int(os.environ["RANK"]) % torch.cuda.device_count()
    return 0


def reduce_dict(input_dict, average: bool = True):
    if get_world_size() < 2:
        return input_dict
    with torch.no_grad():
        keys = sorted(input_dict.keys())
        vals = torch.stack([input_dict[k] for k in keys])
        dist.all_reduce(vals)
        if average:
            vals /= get_world_size()
        return dict(zip(keys, vals))


def broadcast_scalar(value, src: int = 0):
    if get_world_size() == 1:
        return value
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    t = torch.tensor([float(value)], device=dev)
    dist.broadcast(t, src=src)
    return t.item()


def allreduce_mean_(t):
    """In-place mean over ranks of a flat tensor (the gradient arenas).  NCCL averages inside the collective
    (ReduceOp.AVG over NVLink / NVSwitch); gloo (CPU tests) sums and divides."""
    world = get_world_size()
    if world < 2:
        return t
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t)
        t.div_(world)
    return t


def synchronize() -> None:
    if get_world_size() > 1:
        dist.barrier()


def wrap_ddp(model, local_rank=None):
    """Gradient-bucket all-reduce only (the reference's DDP wrap, train.py:167-179, minus the
    per-forward buffer broadcast: BatchNorm statistics are per-rank by design when SyncBN is off)."""
    from torch.nn.parallel import DistributedDataParallel as DDP

    if not is_dist_available_and_initialized():
        return model
    kw = dict(find_unused_parameters=False, broadcast_buffers=False, gradient_as_bucket_view=True,
              bucket_cap_mb=25)
    if torch.cuda.is_available():
        lr = get_local_rank() if local_rank is None else local_rank
        return DDP(model, device_ids=[lr], output_device=lr, **kw)
    return DDP(model, **kw)
