"""One process per GPU; the read stream is sharded per rank; the only collective is the end-of-run
count all-reduce.

The reference runs inference as a multi-process work queue with one model replica per GPU and no
collectives (call_modifications.py:562-578); each worker logs its own counters at exit
(call_modifications.py:405-406,456).  Here ranks are launched by torchrun, hole-batches are assigned
round-robin by index (``batch_idx % world == rank``) or, for synthetic runs, as contiguous ranges, and the
counters {sites_called, model_batches, reads_written, reads_with_MM} are summed with one all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _dev():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def barrier():
    if dist.is_initialized():
        dist.barrier()


def allreduce_max(x):
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_counts(counts):
    """SUM all-reduce of the int64 run counters.  Returns a list of python ints."""
    if not dist.is_initialized():
        return [int(c) for c in counts]
    t = torch.tensor([int(c) for c in counts], dtype=torch.int64, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def shard_range(n, rank, world):
    """Contiguous [start, end) of n units for this rank (synthetic runs)."""
    per = (n + world - 1) // world
    s = min(n, rank * per)
    return s, min(n, s + per)


def owns_holebatch(batch_idx, rank, world):
    """Round-robin hole-batch ownership for BAM input (hole-batch = --holes_batch reads)."""
    return batch_idx % world == rank


def finalize():
    if dist.is_initialized():
        dist.destroy_process_group()
