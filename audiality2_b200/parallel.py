"""Multi-GPU host logic: one process per GPU, voices sharded across ranks.

Leaf voices are independent and every cross-voice coupling is an integer `+=`
into a parent bus (panmix.c:104-105, wtosc.c:229), so any partition of the
voices is legal; the only exchange step is the sum of the stereo ROOT bus
before the first truncating stage, the root panmix (src/audiality2.c:271-280,
SURVEY.md 8(e)).

Two ways to do that exchange:

* `connect_engines` (the product path): every rank's engine maps every other
  rank's "symmetric" buffer (CUDA IPC over NVLink / NVSwitch) and the LAST CTA
  of the render kernel pushes its raw root bus into all peers, waits for the
  world's flags, sums and runs the root stage - no collective call, no extra
  launch (csrc/a2cu_kernels.cuh xchg_root_bus). torch.distributed is only the
  plumbing that swaps the 64-byte buffer handles once.
* `reduce_root_bus`: the plain library collective (NCCL on GPUs, gloo in the CPU
  tests) on a raw root bus produced with `set_post_root_stage(False)`, followed
  by `apply_root_stage` - kept as the baseline the fused path is measured
  against (bench.py --exchange nccl) and for hosts without peer access.
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


def shard_counts(n, world_size):
    return [hi - lo for lo, hi in (shard_range(n, world_size, r) for r in range(world_size))]


def reduce_root_bus(bus):
    """In-place int32 sum of the raw root bus over all ranks (NCCL on GPUs,
    gloo in the CPU tests). No-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bus, op=dist.ReduceOp.SUM)
    return bus


def exchange_handles(handle, device=None):
    """All-gather one 64-byte buffer handle per rank over the default process
    group; returns the list in rank order. Works on gloo (CPU tensors) and on
    nccl (the bytes ride in a CUDA tensor on `device`)."""
    import torch
    import torch.distributed as dist
    n = dist.get_world_size()
    backend = dist.get_backend()
    t = torch.tensor(list(handle), dtype=torch.uint8)
    if backend == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    out = [torch.empty_like(t) for _ in range(n)]
    dist.all_gather(out, t)
    return [bytes(x.cpu().tolist()) for x in out]


def connect_engines(engine, max_frames, timeout_ms=0, device=None):
    """Wire `engine` (this rank's) to all other ranks' engines for the in-kernel
    root-bus exchange. Collective: every rank must call it. Returns world size."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return 1
    n, r = dist.get_world_size(), dist.get_rank()
    handle = engine.xchg_create(r, n, max_frames, timeout_ms)
    engine.xchg_connect_ipc(exchange_handles(handle, device))
    dist.barrier()          # nobody pushes before every rank has mapped every buffer
    return n
