"""Per-role busy cycles inside render_split (a2cu_split_profile) for the cfg2 bank.

    python profiles/role_profile.py [voices]
"""
import sys
sys.path.insert(0, '.')
from audiality2_b200 import engine as eng
from audiality2_b200.workloads import setup_cfg2

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
e = eng.Engine(48000, 2)
bank, b = setup_cfg2(e, V)
e.set_timing(True)
for i in range(5):
    e.write_all(bank, 0, 2, [b['amp'] // (1 + i % 2)], dur=960 << 8)
    e.run(960, 64)
e.split_profile(True, False)
N = 20
ms = []
marks = []
for i in range(N):
    e.write_all(bank, 0, 2, [b['amp'] // (1 + i % 2)], dur=960 << 8)
    e.split_trace_reset()
    e.run(960, 64)
    ms.append(e.last_render_ms())
    gg = e.split_trace()[4, 56:60, 0].astype('int64')
    marks.append(((gg[1] - gg[0]) / 1e3, (gg[2] - gg[0]) / 1e3, (gg[3] - gg[0]) / 1e3, 1e3 * ms[-1]))
tr = e.split_trace()
p = e.split_profile(False, True)
sets = (V + 31) // 32
frag = 15
print('voices %d: kernel %.1f us (min %.1f), split launches %d' % (V, 1e3 * sum(ms) / N, 1e3 * min(ms), e.split_launches))
names = ['control', 'serial compute', 'stage A (one helper)', 'stage C (one helper)', 'serial wait']
for n, v in zip(names, p[:5]):
    print('%-24s %8.0f cycles per fragment and voice set' % (n, v / (N * sets * frag)))
print('fragments counted %d' % p[5])
ncta = max(1, p[5] // (N * frag)) if p[5] else sets
print('per CTA and launch: prologue %.0f cycles, pipeline + state store %.0f cycles' % (
    p[6] / (N * ncta), p[7] / (N * ncta)))

print('timeline of CTA 0, set 0 (cycles since pipeline start): begin-end per fragment')
for r, nm in enumerate(['control', 'serial', 'A helper0', 'C helper0', 'A helperN', 'C helperN']):
    print('%-10s' % nm, ' '.join('%d-%d' % (tr[r, f, 0], tr[r, f, 1]) for f in range(frag)))
g = tr[4, 56:60, 0].astype('int64')
print('grid wall clock (globaltimer): first CTA entry -> last pipeline end %.1f us, -> last CTA done %.1f us, '
      '-> fused tail done %.1f us; CUDA-event span of that launch %.1f us' % (
          (g[1] - g[0]) / 1e3, (g[2] - g[0]) / 1e3, (g[3] - g[0]) / 1e3, 1e3 * ms[-1]))
print('per launch (us): pipeline end, last CTA done, fused tail done, CUDA-event span')
for m in marks:
    print('   %.1f  %.1f  %.1f  %.1f' % m)
