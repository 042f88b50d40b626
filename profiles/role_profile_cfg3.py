"""Per-role busy cycles of render_split for the cfg3 bank (8 x wtosc + panmix)."""
import sys
sys.path.insert(0, '.')
from audiality2_b200 import engine as eng
from audiality2_b200.workloads import setup_cfg3

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
e = eng.Engine(48000, 2)
banks, _ = setup_cfg3(e, V)
e.set_timing(True)
for i in range(3):
    e.run(1024, 256)
e.split_profile(True, False)
N = 4
ms = []
for i in range(N):
    e.run(1024, 256)
    ms.append(e.last_render_ms())
tr = e.split_trace()
p = e.split_profile(False, True)
sets = (V + 31) // 32
frag = 16
print('voices %d: kernel %.1f us (min %.1f), split launches %d' % (V, 1e3 * sum(ms) / N, 1e3 * min(ms), e.split_launches))
for n, v in zip(['control', 'serial compute', 'stage A (one helper)', 'stage C (one helper)', 'serial wait'], p[:5]):
    print('%-24s %8.0f cycles per fragment and voice set' % (n, v / (N * sets * frag)))
print('per CTA and launch: prologue %.0f cycles, pipeline + state store %.0f cycles' % (p[6] / (N * sets), p[7] / (N * sets)))
for r, nm in enumerate(['control', 'serial', 'A helper0', 'C helper0', 'A helperN', 'C helperN']):
    print('%-10s' % nm, ' '.join('%d-%d' % (tr[r, f, 0], tr[r, f, 1]) for f in range(frag)))
