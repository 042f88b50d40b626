"""CPU tests of the N>1 host logic (gloo, world_size 2): sharding voices across
ranks and summing the integer root bus reproduces the single-process render bit
for bit. The per-rank renderer here is the oracle port (no GPU in this suite)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiality2_b200.parallel import exchange_handles, reduce_root_bus, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 64, 4096, 4097):
        for w in (1, 2, 3, 8):
            cuts = [shard_range(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cases import bank
    from scenarios import run_oracle
    scn = bank(48, frames=256)
    lo, hi = shard_range(len(scn.voices), world, rank)
    scn.voices = scn.voices[lo:hi]
    part = torch.from_numpy(run_oracle(scn))
    reduce_root_bus(part)
    # the plumbing of the in-kernel exchange: 64-byte buffer handles, all-gathered in rank order
    handles = exchange_handles(bytes([rank * 16 + (i % 16) for i in range(64)]))
    assert len(handles) == world
    for r, h in enumerate(handles):
        assert h == bytes([r * 16 + (i % 16) for i in range(64)])
    if rank == 0:
        q.put(part.numpy())
    dist.destroy_process_group()


def test_two_rank_bus_reduce_equals_single_render():
    from cases import bank
    from scenarios import run_oracle
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = run_oracle(bank(48, frames=256))
    assert np.array_equal(out, whole)
