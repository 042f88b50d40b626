"""Kernel-span throughput of BASELINE configs 3 and 4 (bank mode, static banks),
for profiles/README.md. Not the bench contract line (that is cfg2, bench.py)."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from audiality2_b200 import engine as eng
from audiality2_b200.chains import autowire
from audiality2_b200.workloads import fx
from cases import _FM_SETTINGS


def run(e, frames, buffer, steps):
    e.set_timing(True)
    for _ in range(3):
        e.run(frames, buffer)
    ms = 0.0; t = 0.0
    for _ in range(steps):
        t0 = time.perf_counter(); e.run(frames, buffer); t += time.perf_counter() - t0
        ms += e.last_render_ms() + e.last_mix_ms()
    return ms / steps, t / steps * 1e3


def cfg3(V=65536):
    e = eng.Engine(48000, 2); w = e.builtin_wave("sine")
    r = np.random.RandomState(3)
    pitch = r.randint(-2 * 65536, 2 * 65536, size=V).astype(np.int32)
    bank = e.new_bank(autowire(["wtosc"] * 8 + ["panmix"]), V)
    for k in range(8):
        e.write_all(bank, k, 1, pitch + fx(np.log2(k + 1))); e.write_all(bank, k, 2, [fx(0.00002 / (k + 1))])
        e.write_all(bank, k, 0, [w << 16])
    e.write_all(bank, 8, 1, r.randint(-65536, 65536, size=V).astype(np.int32))
    for split in (True, False):
        e.set_split(split)
        k_ms, h_ms = run(e, 1024, 256, 10)
        print("cfg3 %d voices x 8 wtosc, 256-frame buffers, 1024 frames/step, %s: kernels %.3f ms -> %.1f G voice-samples/s (%.1f G osc-samples/s), a2cu_run %.3f ms" % (
            V, "render_split" if split else "render_bank", k_ms, V * 1024 / k_ms / 1e6, 8 * V * 1024 / k_ms / 1e6, h_ms))
    e.close()


def cfg4(V=32768):
    e = eng.Engine(48000, 2)
    r = np.random.RandomState(11)
    for ki, kind in enumerate(["fm3", "fm3p", "fm2r", "fm4r"]):
        n = V // 4
        bank = e.new_bank(autowire([kind, "panmix"]), n)
        st = _FM_SETTINGS[kind]
        e.write_all(bank, 0, 1, r.randint(-2 * 65536, 2 * 65536, size=n).astype(np.int32))
        e.write_all(bank, 0, 2, [fx(0.0005)]); e.write_all(bank, 0, 3, [fx(st[0][1])])
        for op in range(1, len(st)):
            p, a, fb = st[op]
            e.write_all(bank, 0, 1 + 3 * op, [fx(p)]); e.write_all(bank, 0, 2 + 3 * op, [fx(a)])
            e.write_all(bank, 0, 3 + 3 * op, [fx(fb)])
        e.write_all(bank, 1, 1, r.randint(-65536, 65536, size=n).astype(np.int32))
    k_ms, h_ms = run(e, 960, 64, 10)
    print("cfg4 shard %d FM voices (fm3/fm3p/fm2r/fm4r), 64-frame blocks, 960 frames/step: kernels %.3f ms -> %.1f G voice-samples/s, a2cu_run %.3f ms" % (
        V, k_ms, V * 960 / k_ms / 1e6, h_ms))
    e.close()


cfg3(); cfg4(); cfg4(262144)
