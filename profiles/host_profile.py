"""Host-side breakdown of one bench step on the GPU box: bulk write, a2cu_submit
(event staging + H2D + launches + D2H queued) and a2cu_collect, two windows in
flight like bench.py; A2CU_STATS=1 adds the engine's own per-window split."""
import os, sys, time
os.environ["A2CU_STATS"] = "1"
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from audiality2_b200 import engine as eng
from audiality2_b200.workloads import cfg2_bank
from audiality2_b200.chains import autowire
e = eng.Engine(48000, 2); b = cfg2_bank(4096); w = e.builtin_wave('saw')
bank = e.new_bank(autowire(list(b['kinds'])), 4096)
e.write_all(bank, 0, 0, [w << 16]); e.write_all(bank, 0, 1, b['pitch']); e.write_all(bank, 0, 2, [b['amp']])
e.write_all(bank, 1, 0, b['cutoff']); e.write_all(bank, 1, 1, [b['q']]); e.write_all(bank, 2, 1, b['pan'])
e.set_timing(True)
L = e.L
amp = [np.array([b['amp']], dtype=np.int32), np.array([b['amp'] // 2], dtype=np.int32)]
out = np.empty((960, 2), dtype=np.int32)
for i in range(20):
    e.write_all(bank, 0, 2, amp[i & 1], dur=960 << 8); e.run(960, 64)
N = 400
tw = ts = tc = 0.0; k = 0.0
pend = []
T0 = time.perf_counter()
for i in range(N):
    t0 = time.perf_counter()
    L.a2cu_bank_write_all(e.h, bank, 0, 2, amp[i & 1].ctypes.data, 0, L.a2cu_now(e.h), 960 << 8)
    t1 = time.perf_counter()
    pend.append(e.submit(960, 64))
    t2 = time.perf_counter()
    if len(pend) > 2:
        e.collect(pend.pop(0), out); k += e.last_render_ms() + e.last_mix_ms()
    t3 = time.perf_counter()
    tw += t1 - t0; ts += t2 - t1; tc += t3 - t2
while pend:
    e.collect(pend.pop(0), out)
T1 = time.perf_counter()
print("per step: write_all %.1f us  submit %.1f us  collect(wait) %.1f us  loop total %.1f us  (kernels %.1f us)" % (
    tw / N * 1e6, ts / N * 1e6, tc / N * 1e6, (T1 - T0) / N * 1e6, k / (N - 2) * 1e3))
e.close()
