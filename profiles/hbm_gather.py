"""The genuinely HBM-bound case of the wtosc gather (SURVEY.md 8(d) "honest caveat"): large SAMPLED
waves. Workload = audiality2_b200.workloads.setup_gather (the same function bench.py's
configs.gather uses): 12 looped sampled waves x 16.8 M samples (403 MB, 3 x L2) played at 64 wave
samples per output frame, 2 Hermite taps per output sample, each in its own 32-byte sector.

    python profiles/hbm_gather.py [voices] [windows] [wave samples per frame]
"""
import json
import sys

sys.path.insert(0, '.')
import numpy as np
from audiality2_b200 import engine as eng
from audiality2_b200 import workloads as wl

V = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
FRAMES = 256
try:
    peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    src = 'measured (MEASURED_PEAKS.json)'
except Exception:
    peak, src = 6650.0, 'fallback'
e = eng.Engine(48000, 2)
banks, info = wl.setup_gather(e, V, samples_per_frame=S)
e.set_timing(True)
out = e.run(FRAMES, 64)
ms = []
for i in range(N):
    e.run(FRAMES, 64)
    ms.append(e.last_render_ms())
ms = sorted(ms)[len(ms) // 2]
vs = V * FRAMES
print(json.dumps({
    "workload": "%d voices, %.0f MB of sampled waves, %d wave samples per frame, %d-frame windows" % (
        V, info["wave_bytes"] / 1e6, S, FRAMES),
    "kernel": ("render_split" if e.split_launches else "render_bank") + "<" + e.bank_kernel_name(banks[0]) + ">",
    "ms_per_window": ms, "voice_samples_per_s": vs / ms * 1e3,
    "algorithmic_GBps": vs * 64 / ms / 1e6, "hbm_peak_GBps": peak, "peak_source": src,
    "frac": vs * 64 / ms / 1e6 / peak, "peak_abs": int(np.abs(out).max())}))
e.close()
