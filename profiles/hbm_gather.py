"""The genuinely HBM-bound case of the wtosc gather (SURVEY 8(d) "honest caveat"): large SAMPLED waves.

The benchmark configs play a 2048-point builtin wave that lives in shared memory, so their gather
never touches HBM. Here NW non-mipmapped sampled waves of 16 M samples each (the reference's limit is
2^24 - 133 frames, wtosc.c:55) - together several times the 126 MB L2 - are played by V voices
{wtosc; panmix} at S wave samples per output frame from random start phases. With S = 64 every output
sample interpolates at two places (a2_Hermite at ph and ph + dph/2, wtosc.c:226-228) 64 bytes apart
or more: each touches its own 32-byte sector(s), so the algorithmic HBM traffic is 2 sectors = 64 B
per voice-sample.

    python profiles/hbm_gather.py [voices] [waves] [S]
"""
import sys
import time

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from audiality2_b200 import engine as eng
from audiality2_b200.chains import autowire

V = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 12
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
LEN = (1 << 24) - 256
FRAMES = 256

e = eng.Engine(48000, 2)
rng = np.random.RandomState(1)
base = rng.randint(-20000, 20000, size=LEN).astype(np.int16)
# dphase at pitch 0 (incl. basepitch): read it back through one probe voice's math instead of
# re-deriving it: period P gives dph = dphase * P per frame; we want dph = S << 24
dphase0 = 261.626 / 48000.0 * (1 << 24)
P = int(round(S * (1 << 24) / dphase0))
waves = []
t0 = time.time()
for w in range(NW):
    waves.append(e.upload_wave(2, P, 0x100, np.roll(base, 7919 * w)))
print("uploaded %d waves x %d samples (%.0f MB int16), period %d, %.1f s" % (
    NW, LEN, NW * LEN * 2 / 1e6, P, time.time() - t0), flush=True)
bank = e.new_bank(autowire(["wtosc", "panmix"]), V)
e.write_all(bank, 0, 0, (np.array([waves[v % NW] for v in range(V)], dtype=np.int64) << 16).astype(np.int32))
e.write_all(bank, 0, 1, [0])
e.write_all(bank, 0, 2, [65])
periods = LEN // P
e.write_all(bank, 0, 3, (rng.randint(0, periods - 2, size=V).astype(np.int64) << 16).astype(np.int32))
e.write_all(bank, 1, 1, rng.randint(-65536, 65536, size=V).astype(np.int32))
e.set_timing(True)
t0 = time.time()
out = e.run(FRAMES, 64)
print("first window %.2f s (wave pool upload), peak %d" % (time.time() - t0, int(np.abs(out).max())), flush=True)
ms = []
for i in range(6):
    e.run(FRAMES, 64)
    ms.append(e.last_render_ms())
ms = sorted(ms)[len(ms) // 2]
vs = V * FRAMES
print("kernel %s: %.3f ms per %d-frame window, %.1f G voice-samples/s" % (
    e.bank_kernel_name(bank), ms, FRAMES, vs / ms / 1e6))
print("algorithmic gather traffic 64 B per voice-sample: %.0f GB/s (of 6650 GB/s HBM fallback peak: %.1f%%)" % (
    vs * 64 / ms / 1e6, 100 * vs * 64 / ms / 1e6 / 6650))
e.close()
