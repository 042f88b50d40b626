"""torchrun worker of tests/test_multigpu.py::test_two_process_exchange_over_ipc.

One process per GPU. Every rank renders its round-robin share of a bank's voices twice -
once with the in-kernel peer-memory exchange (a2cu_xchg_*, handles swapped over
torch.distributed), once with the cut path + NCCL all-reduce - and compares both with ONE
engine rendering all voices on its own GPU. Prints MGPU_OK when every rank agrees.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import bank  # noqa: E402
from scenarios import _build_cuda_shard, run_cuda  # noqa: E402
from audiality2_b200 import engine as eng  # noqa: E402
from audiality2_b200.parallel import connect_engines, reduce_root_bus  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    scn = bank(384, frames=1600)
    scn.root_writes = [(300 << 8, 0, 45000, 400 << 8), (1000 << 8, 1, -25000, 0)]
    window = scn.buffer * 5
    mine = range(rank, len(scn.voices), world)

    # single engine, all voices, on this rank's GPU
    os.environ["A2CU_DEVICE"] = str(local)
    whole_engine = eng.Engine(scn.samplerate, scn.channels, device=local)
    _build_cuda_shard(scn, whole_engine, range(len(scn.voices)))
    whole = whole_engine.run(scn.frames, scn.buffer)
    whole_engine.close()

    # fused: exchange inside the render kernel
    e = eng.Engine(scn.samplerate, scn.channels, device=local)
    st = torch.cuda.Stream()
    e.set_stream(st.cuda_stream)
    _build_cuda_shard(scn, e, mine)
    n = connect_engines(e, window, timeout_ms=20000, device=dev)
    assert n == world
    parts, tickets, done = [], [], 0
    while done < scn.frames:
        k = min(window, scn.frames - done)
        tickets.append(e.submit(k, scn.buffer))
        done += k
        if len(tickets) > 2:
            parts.append(e.collect(tickets.pop(0)))
    while tickets:
        parts.append(e.collect(tickets.pop(0)))
    fused = np.concatenate(parts, axis=0)
    split_launches = e.split_launches
    e.close()

    # baseline: cut path + NCCL all-reduce + root stage
    e = eng.Engine(scn.samplerate, scn.channels, device=local)
    e.set_stream(st.cuda_stream)
    _build_cuda_shard(scn, e, mine)
    e.set_post_root_stage(False)
    parts, done = [], 0
    with torch.cuda.stream(st):
        while done < scn.frames:
            k = min(window, scn.frames - done)
            bus = torch.zeros((k, 2), dtype=torch.int32, device=dev)
            master = torch.zeros((k, scn.channels), dtype=torch.int32, device=dev)
            e.run_async(k, scn.buffer, bus.data_ptr())
            reduce_root_bus(bus)
            e.apply_root_stage(bus.data_ptr(), master.data_ptr(), k, scn.buffer)
            st.synchronize()
            parts.append(master.cpu().numpy())
            done += k
    nccl = np.concatenate(parts, axis=0)
    e.close()

    ok = bool(np.array_equal(fused, whole) and np.array_equal(nccl, whole) and np.abs(whole).max() > 1000)
    if not ok:
        print("rank %d: fused==whole %s, nccl==whole %s, peak %d" % (
            rank, np.array_equal(fused, whole), np.array_equal(nccl, whole), int(np.abs(whole).max())),
            flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("MGPU_OK world %d, %d windows, split launches %d" % (world, (scn.frames + window - 1) // window,
                                                                split_launches), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
