"""Where the tail of a sharded render_split launch spends its time (a2cu_split_trace, role 5
fragments 60..62): two engines of one process on cuda:0 / cuda:1 (or both on cuda:0), cfg2 banks,
pipelined submit/collect so the lagged exchange is active.

    python profiles/xchg_tail.py [ndev] [engines]
"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from audiality2_b200 import engine as eng
from audiality2_b200.workloads import setup_cfg2

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else min(2, torch.cuda.device_count())
NE = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W = 960
engines, banks, amps, streams = [], [], [], []
for r in range(NE):
    d = r % ndev
    torch.cuda.set_device(d)
    e = eng.Engine(48000, 2, device=d)
    st = torch.cuda.Stream(device=d)
    e.set_stream(st.cuda_stream)
    e.set_timing(True)
    bank, b = setup_cfg2(e, 4096, seed=324357 + r)
    engines.append(e); banks.append(bank); amps.append(b['amp']); streams.append(st)
for r, e in enumerate(engines):
    e.xchg_create(r, NE, W, timeout_ms=5000)
for e in engines:
    e.xchg_connect_local(engines)


def run(n, profile=False):
    tick = [[] for _ in engines]
    ms = []
    for i in range(n):
        for r, e in enumerate(engines):
            e.write_all(banks[r], 0, 2, [amps[r] // (1 + i % 2)], dur=W << 8)
            tick[r].append(e.submit(W, 64))
        if len(tick[0]) > 2:
            for r, e in enumerate(engines):
                e.collect(tick[r].pop(0))
                ms.append(e.last_render_ms() + e.last_mix_ms())
    while tick[0]:
        for r, e in enumerate(engines):
            e.collect(tick[r].pop(0))
    return ms


run(20)
for e in engines:
    e.split_profile(True, False)
ms = run(40)
print('%d engines on %d device(s): kernel span %.1f us (median), %.1f (min)' % (NE, ndev, 1e3 * np.median(ms), 1e3 * min(ms)))
for r, e in enumerate(engines):
    tr = e.split_trace()
    t = tr[5, 60:63, 0]
    ser_end = tr[1, 14, 1]
    print('engine %d: last recurrence fragment ends %d, tail starts %d, previous window finished +%d, published +%d cycles'
          % (r, ser_end, t[0], t[1] - t[0], t[2] - t[1]))
    e.split_profile(False, False)
for e in engines:
    e.close()
