"""Host-side breakdown of one bench step (write_all + a2cu_run) on the GPU box."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from audiality2_b200 import engine as eng
from audiality2_b200.workloads import cfg2_bank
from scenarios import autowire
e = eng.Engine(48000, 2); b = cfg2_bank(4096); w = e.builtin_wave('saw')
bank = e.new_bank(autowire(list(b['kinds'])), 4096)
e.write_all(bank, 0, 0, [w << 16]); e.write_all(bank, 0, 1, b['pitch']); e.write_all(bank, 0, 2, [b['amp']])
e.write_all(bank, 1, 0, b['cutoff']); e.write_all(bank, 1, 1, [b['q']]); e.write_all(bank, 2, 1, b['pan'])
e.set_timing(True)
for i in range(20):
    e.write_all(bank, 0, 2, [b['amp'] // (1 + i % 2)], dur=960 << 8); e.run(960, 64)
N = 200
tw = tr = 0.0; k = 0.0
for i in range(N):
    t0 = time.perf_counter()
    e.write_all(bank, 0, 2, [b['amp'] // (1 + i % 2)], dur=960 << 8)
    t1 = time.perf_counter()
    e.run(960, 64)
    t2 = time.perf_counter()
    tw += t1 - t0; tr += t2 - t1; k += e.last_render_ms() + e.last_mix_ms()
print("write_all %.1f us  run %.1f us  (kernels %.1f us)  per step" % (tw / N * 1e6, tr / N * 1e6, k / N * 1e3))
# run without events (static), for the fixed launch+sync+copy cost
ts = 0.0
for i in range(N):
    t1 = time.perf_counter(); e.run(960, 64); ts += time.perf_counter() - t1
print("run without events %.1f us (kernels %.1f us)" % (ts / N * 1e6, e.last_render_ms() * 1e3 + e.last_mix_ms() * 1e3))
