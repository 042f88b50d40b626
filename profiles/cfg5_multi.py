"""BASELINE config 5 on several GPUs: k2trance.a2s `Song` x N, the Song instances dealt round-robin
over `world` host processes (one reference host + unit plug-in + engine per GPU, A2CU_DEVICE = rank),
int32 outputs summed (the cut is the root bus; the root panmix is the identity here,
so summing master blocks is the same integer sum).

One engine state's VM `rand` and noise oscillators share one LCG in tree-walk order
(core.c:1401-1409), so a sharded render is NOT comparable with the single-state render of all
copies; the oracle for a sharded run is the reference itself run on the same shards - which is also
the reference's own way to use several cores (independent states, audiality2.h.cmake:163-166).

    python profiles/cfg5_multi.py [copies] [frames] [world]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import a2oracle as ao  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 88200
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
song = os.path.join(ao.REF_DIR, "songs", "benchmark", "k2trance.a2s")


def render(binary, nshards, gpus):
    """Run `nshards` processes concurrently; returns (summed int32 output, max a2_Run seconds, wall)."""
    tmp = tempfile.mkdtemp(prefix="a2cfg5_")
    procs = []
    t0 = time.perf_counter()
    for r in range(nshards):
        env = dict(os.environ)
        if gpus:
            env["A2CU_DEVICE"] = str(r % gpus)
        out = os.path.join(tmp, "s%d.raw" % r)
        procs.append((out, subprocess.Popen(
            [os.path.join(ao.REF_DIR, binary), "-r", "44100", "-b", "500", "-n", str(frames), "-x", str(copies),
             "-X", "%d/%d" % (r, nshards), "-p", "Song", "-o", out, os.path.basename(song)],
            cwd=os.path.dirname(song), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)))
    total, secs, voices = None, [], 0
    for out, p in procs:
        so, se = p.communicate()
        if p.returncode:
            raise RuntimeError(se[-2000:])
        info = json.loads(so.strip().splitlines()[-1])
        secs.append(info["seconds"])
        voices += info["active_voices"]
        a = np.fromfile(out, dtype="<i4").reshape(-1, 2).astype(np.int64)
        total = a if total is None else total + a
    wall = time.perf_counter() - t0
    total = ((total + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)
    return total, max(secs), wall, voices


import torch  # noqa: E402  (device count only)
ngpu = torch.cuda.device_count()
world = min(world, max(ngpu, 1))
ref1, ref1_s, _, v1 = render("a2render", 1, 0)                      # the reference, one state
refN, refN_s, refN_wall, vN = render("a2render", world, 0)          # the reference, `world` states on host cores
outN, outN_s, outN_wall, vo = render("a2render_cuda", world, ngpu)  # drop-in, one process per GPU
out1, out1_s, _, _ = render("a2render_cuda", 1, 1)                  # drop-in, one GPU
bad = np.nonzero((outN != refN).any(axis=1))[0]
print(json.dumps({
    "workload": "cfg5: k2trance.a2s Song x %d, 44.1 kHz, buffer 500, %d frames, Song instances dealt over %d "
                "processes / GPUs" % (copies, frames, world),
    "gpus": ngpu, "host_threads": os.cpu_count(),
    "sharded_bit_exact_vs_sharded_reference": bool(len(bad) == 0), "first_diff": int(bad[0]) if len(bad) else None,
    "single_bit_exact_vs_single_reference": bool(np.array_equal(out1, ref1)),
    "active_voices_end": vo, "ref_active_voices_end": vN,
    "reference_1_state_s": ref1_s, "reference_%d_states_s" % world: refN_s,
    "dropin_1_gpu_s": out1_s, "dropin_%d_gpus_s" % world: outN_s,
    "speedup_vs_1_state": ref1_s / outN_s, "speedup_vs_same_number_of_states": refN_s / outN_s,
    "peak": int(np.abs(refN).max())}))
