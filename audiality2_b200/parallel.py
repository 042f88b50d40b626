"""Multi-GPU host logic: voices shard across ranks with no data-path collective
except ONE integer sum of the stereo root bus per window.

The cut is the bus that feeds the first truncating stage - the root scratch bus
before the root panmix (src/audiality2.c:271-280). Integer addition is
associative and commutative, so any partition and any reduction order gives the
bit-identical bus; everything above the cut runs once, after the reduce.
"""
import os


def world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), \
        int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n, world_size, rank):
    """Contiguous, balanced [lo, hi) of n voices for `rank`."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_root_bus(bus):
    """In-place int32 sum of the raw root bus over all ranks (NCCL on GPUs,
    gloo in the CPU tests). No-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bus, op=dist.ReduceOp.SUM)
    return bus
