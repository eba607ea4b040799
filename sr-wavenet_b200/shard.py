"""Utterance-batch sharding across the GPUs of one box.

Teacher scoring, student synthesis and autoregressive generation treat every utterance
independently (no cross-batch op in model.py:158-200 / 415-535), so ranks take contiguous slices
of the batch and no data-path collective is needed; the only cross-rank traffic is the scalar
reduction of timings / log-likelihood sums."""
import os

import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """Contiguous [start, end) slice of `batch` utterances for `rank`; sizes differ by at most 1."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """One process per GPU under torchrun; returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def reduce_scalars(values, op, device=None):
    """All-reduce a few Python floats (timings: MAX; processed units, NLL sums: SUM)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.tolist()


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


# ---- data-parallel distillation step (the one path with a collective) ---------------------------
def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def global_batch(local_batch, device=None):
    """Sum of the ranks' local batch sizes (model.py:379 divides the loss by the batch: under data parallelism that is
    the GLOBAL batch, and ranks need not hold equal shares)."""
    total, = reduce_scalars([float(local_batch)], "sum", device)
    return int(round(total))


def broadcast_from_rank0(tensors):
    """Make every replica start from rank 0's state (weights, Adam moments, step count): replicas that apply a shared
    gradient to different parameters drift apart silently."""
    if is_distributed():
        for t in tensors:
            dist.broadcast(t, src=0)


def all_reduce_sum(bucket):
    """ONE all-reduce for the whole step: bucket = [flat gradient | loss | power_loss]."""
    if is_distributed():
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    return bucket
