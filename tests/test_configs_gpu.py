"""GPU parity at BASELINE.json's full configuration sizes, driven through the
bulk C-ABI calls (a2cu_bank_write_all) against the oracle port:

  cfg 3  65 536 voices, 8 x wtosc + panmix per voice, sine, 256-frame buffers
  cfg 4  the per-GPU shard (32 768 voices) of the 262 144-voice FM bank cycling
         fm3 / fm3p / fm2r / fm4r + panmix, 64-frame blocks
"""
import numpy as np
import pytest

from cases import _FM_SETTINGS
from scenarios import autowire, fx
from oracle import a2oracle as ao

pytestmark = pytest.mark.gpu


def _diff(a, b):
    bad = np.nonzero((a != b).any(axis=1))[0]
    return "first diff at frame %d (%d frames differ, max abs %d)" % (
        bad[0], len(bad), np.abs(a.astype(np.int64) - b).max())


def test_cfg3_additive_65536_voices():
    from audiality2_b200 import engine as eng
    V, NOSC, frames, buffer = 65536, 8, 256, 256
    r = np.random.RandomState(3)
    pitch = r.randint(-2 * 65536, 2 * 65536, size=V).astype(np.int32)
    pan = r.randint(-65536, 65536, size=V).astype(np.int32)
    kinds = ["wtosc"] * NOSC + ["panmix"]
    chain = autowire(kinds)

    e = eng.Engine(48000, 2)
    w = e.builtin_wave("sine")
    bank = e.new_bank(chain, V)
    o = ao.Oracle(48000, 2)
    ow = o.builtin_wave("sine")
    for _ in range(V):
        o.new_voice(chain)
    for k in range(NOSC):
        pk = pitch + fx(np.log2(k + 1))
        amp = fx(0.00002 / (k + 1))
        e.write_all(bank, k, 1, pk)
        e.write_all(bank, k, 2, [amp])
        e.write_all(bank, k, 0, [w << 16])
        o.write_all(0, V, k, 1, pk)
        o.write_all(0, V, k, 2, [amp])
        o.write_all(0, V, k, 0, [ow << 16])
    e.write_all(bank, NOSC, 1, pan)
    o.write_all(0, V, NOSC, 1, pan)
    out = e.run(frames, buffer)
    assert e.split_launches > 0          # warp-specialised kernel, 8 oscillators
    ref = o.render(np.zeros(0, dtype=ao.EVENT_DTYPE), frames, buffer)
    e.close()
    o.close()
    assert np.abs(ref).max() > 10000
    assert np.array_equal(out, ref), _diff(out, ref)


def test_cfg4_fm_shard_32768_voices():
    from audiality2_b200 import engine as eng
    V, frames = 32768, 256
    kinds = ["fm3", "fm3p", "fm2r", "fm4r"]
    r = np.random.RandomState(11)
    pitch = r.randint(-2 * 65536, 2 * 65536, size=V).astype(np.int32)
    pan = r.randint(-65536, 65536, size=V).astype(np.int32)

    e = eng.Engine(48000, 2)
    o = ao.Oracle(48000, 2)
    first = 0
    for ki, kind in enumerate(kinds):
        idx = np.arange(ki, V, 4)
        n = len(idx)
        chain = autowire([kind, "panmix"])
        bank = e.new_bank(chain, n)
        for _ in range(n):
            o.new_voice(chain)
        st = _FM_SETTINGS[kind]

        def both(unit, reg, values, dur=0):
            e.write_all(bank, unit, reg, values, dur=dur)
            o.write_all(first, n, unit, reg, values, 0, dur)

        both(0, 1, pitch[idx])
        both(0, 2, [fx(0.0005 * st[0][0])])
        both(0, 3, [fx(st[0][1])])
        for op in range(1, len(st)):
            p, a, fb = st[op]
            both(0, 1 + 3 * op, [fx(p)])
            both(0, 2 + 3 * op, [fx(a)])
            both(0, 3 + 3 * op, [fx(fb)], dur=100 << 8)     # one ramping feedback
        both(1, 1, pan[idx])
        first += n
    out = e.run(frames, 64)
    ref = o.render(np.zeros(0, dtype=ao.EVENT_DTYPE), frames, 64)
    e.close()
    o.close()
    assert np.abs(ref).max() > 10000
    assert np.array_equal(out, ref), _diff(out, ref)


@pytest.mark.parametrize("wtype,looped", [(2, True), (2, False), (3, True)])
def test_uploaded_sampled_waves_all_gather_modes(wtype, looped):
    """Uploaded (sampled) waves, CUDA vs the oracle port: non-mipmapped waves below and above
    A2_MAXPHINC samples per frame (the per-sample wrapped loop / end check of wtosc.c:301-358, i.e.
    the gather that goes to HBM for large waves, profiles/hbm_gather.py), non-looped waves that
    run out mid-window, and an uploaded mipmapped wave. Raw int16 taps (no coefficient table)
    and the table path are both hit: the second wave is longer than the table limit."""
    from audiality2_b200 import engine as eng
    V, frames = 96, 640
    r = np.random.RandomState(5 + wtype + int(looped))
    small = (r.randint(-30000, 30000, size=3001)).astype(np.int16)          # 3001: not a multiple of 256
    big = (r.randint(-30000, 30000, size=(1 << 20) + 77)).astype(np.int16)   # > 2^20 samples: raw-tap path
    flags = 0x100 if looped else 0
    kinds = ["wtosc", "panmix"]
    chain = autowire(kinds)
    e = eng.Engine(48000, 2)
    o = ao.Oracle(48000, 2)
    waves_e = [e.upload_wave(wtype, 500, flags, small), e.upload_wave(wtype, 40000, flags, big)]
    waves_o = [o.upload_wave(wtype, 500, flags, small), o.upload_wave(wtype, 40000, flags, big)]
    bank = e.new_bank(chain, V)
    for _ in range(V):
        o.new_voice(chain)
    # pitches from -3 to +6.5 octaves: period 500 -> up to ~250 samples per frame, period 40000 -> far
    # beyond A2_MAXPHINC (muted / checked paths)
    pitch = (np.linspace(-3, 6.5, V) * 65536).astype(np.int32)
    phase = r.randint(0, 3 * 65536, size=V).astype(np.int32)
    pan = r.randint(-65536, 65536, size=V).astype(np.int32)
    we = np.array([waves_e[v % 2] << 16 for v in range(V)], dtype=np.int32)
    wo = np.array([waves_o[v % 2] << 16 for v in range(V)], dtype=np.int32)
    e.write_all(bank, 0, 0, we); o.write_all(0, V, 0, 0, wo)
    for reg, vals in ((1, pitch), (2, [fx(0.01)]), (3, phase)):
        e.write_all(bank, 0, reg, vals); o.write_all(0, V, 0, reg, vals)
    e.write_all(bank, 1, 1, pan); o.write_all(0, V, 1, 1, pan)
    out = e.run(frames, 64)
    ref = o.render(np.zeros(0, dtype=ao.EVENT_DTYPE), frames, 64)
    e.close()
    o.close()
    assert np.abs(ref).max() > 10000
    assert np.array_equal(out, ref), _diff(out, ref)


def test_cfg4_full_262144_fm_voices():
    """BASELINE config 4 at its full size on ONE GPU: 262 144 voices cycling the four fmtest4
    instruments, set up by the same function bench.py uses (workloads.setup_cfg4), against the port."""
    from audiality2_b200 import engine as eng
    from audiality2_b200 import workloads as wl
    V, frames = 262144, 128
    e = eng.Engine(48000, 2)
    o = ao.Oracle(48000, 2)
    first = [0, 0, 0, 0]
    created = [False] * 4

    def mirror(ki, chain, n, unit, reg, values, dur):
        if not created[ki]:
            first[ki] = sum(V // 4 for k in range(ki))
            for _ in range(n):
                o.new_voice(chain)
            created[ki] = True
        o.write_all(first[ki], n, unit, reg, values, 0, dur)

    wl.setup_cfg4(e, V, writer=mirror)
    out = e.run(frames, 64)
    ref = o.render(np.zeros(0, dtype=ao.EVENT_DTYPE), frames, 64)
    e.close()
    o.close()
    assert np.abs(ref).max() > 100000
    assert np.array_equal(out, ref), _diff(out, ref)


def test_cfg5_k2trance_x100_dropin_equals_reference():
    """BASELINE config 5 at a tenth of its size: the reference's own k2trance.a2s `Song` started
    100 times, 44.1 kHz, 500-frame buffers (benchmark/benchmark.sh:50) - the unmodified reference
    host on our unit plug-in against the full reference, bit for bit (x1000 is measured, and compared
    the same way, by profiles/cfg5_k2trance.py)."""
    import os
    song = os.path.join(ao.REF_DIR, "songs", "benchmark", "k2trance.a2s")
    harness = os.path.join(ao.REF_DIR, "a2render_cuda")
    if not (os.path.exists(song) and os.path.exists(harness)):
        pytest.skip("reference build / drop-in harness not present")
    kw = dict(samplerate=44100, channels=2, buffer=500, frames=30000, copies=100,
              cwd=os.path.dirname(song))
    ref, ri = ao.ref_render(os.path.basename(song), "Song", **kw)
    out, oi = ao.ref_render(os.path.basename(song), "Song", binary="a2render_cuda", **kw)
    assert ri["rt_error"] == 0 and oi["rt_error"] == 0
    assert oi["active_voices"] == ri["active_voices"] and ri["active_voices"] > 400
    assert np.abs(ref).max() > 1000000
    assert np.array_equal(out, ref), _diff(out, ref)


def test_cfg5_sharded_dropin_equals_sharded_reference():
    """Config 5 sharded: the Song instances dealt over two engine states. One state's
    VM `rand` and noise oscillators share one LCG in tree-walk order, so the oracle of a sharded run
    is the reference run on the same shards (its own way to use several cores); the int32 outputs of
    the shards add up. Both drop-in shards run on cuda:0 here; profiles/cfg5_multi.py puts one process
    on each GPU of the box."""
    import os
    song = os.path.join(ao.REF_DIR, "songs", "benchmark", "k2trance.a2s")
    harness = os.path.join(ao.REF_DIR, "a2render_cuda")
    if not (os.path.exists(song) and os.path.exists(harness)):
        pytest.skip("reference build / drop-in harness not present")
    kw = dict(samplerate=44100, channels=2, buffer=500, frames=20000, copies=40, cwd=os.path.dirname(song))
    ref = sum(ao.ref_render(os.path.basename(song), "Song", shard=(r, 2), **kw)[0].astype(np.int64) for r in range(2))
    out = sum(ao.ref_render(os.path.basename(song), "Song", shard=(r, 2), binary="a2render_cuda", **kw)[0]
              .astype(np.int64) for r in range(2))
    whole = ao.ref_render(os.path.basename(song), "Song", **kw)[0]
    assert np.abs(ref).max() > 1000000
    assert np.array_equal(out, ref), _diff(out, ref)
    assert not np.array_equal(ref, whole.astype(np.int64))      # the shared LCG: sharding is a different song


@pytest.mark.parametrize("kinds,buffer", [
    (["wtosc", "panmix"], 64),
    (["wtosc", "panmix"], 50),
    (["wtosc", "wtosc", "panmix"], 64),
    (["wtosc", "filter12", "panmix"], 64),
])
def test_raw_tap_gather_with_segments_and_chains(kinds, buffer):
    """The raw-tap gather of render_split (stage A with lane = frame, csrc/a2cu_split.cuh) on a
    sampled wave too long for a coefficient table, with everything that cuts a fragment into
    segments: pitch / amplitude ramps and phase writes at times that are no multiple of the
    fragment length, voices below and above A2_MAXPHINC samples per frame (plain and per-sample
    wrapped loop, wtosc.c:200-236 and 301-358), two oscillators per voice, a filter behind the
    oscillator, and a driver buffer that is not a multiple of 64. CUDA vs the oracle port."""
    from scenarios import Scenario, run_cuda, run_oracle
    from cases import W, P, A, PH, PAN
    s = Scenario(48000, 2, buffer, 1500)
    w = s.upload(2, 700, 0x100, (1 << 20) + 5003, 11)
    nosc = kinds.count("wtosc")
    pm = len(kinds) - 1
    for i in range(40):
        p0 = -2.5 + 0.27 * i            # up to +8 octaves: period 700 -> far above A2_MAXPHINC
        steps = []
        for k in range(nosc):
            steps += [("ramp", k, W, w << 16), ("set", k, P, fx(p0 + 0.31 * k)), ("set", k, A, fx(0.08)),
                      ("set", k, PH, fx(0.13 * i + k))]
        if "filter12" in kinds:
            steps += [("set", 1, 0, fx(p0 + 1.0)), ("set", 1, 1, fx(1.5))]
        steps += [("set", pm, PAN, fx(-0.9 + 0.045 * i)), ("d", fx(3.1 + 0.07 * i)),
                  ("ramp", 0, P, fx(p0 + 0.8)), ("ramp", 0, A, fx(0.03)), ("d", fx(9.25)),
                  ("set", 0, PH, fx(0.5)), ("ramp", nosc - 1, A, fx(0.1)), ("d", fx(7.3)),
                  ("ramp", 0, P, fx(p0 - 0.5)), ("d", fx(6))]
        s.add_voice(kinds, steps)
    stats = {}
    out = run_cuda(s, stats=stats)
    ref = run_oracle(s)
    assert stats["split_launches"] > 0
    assert np.abs(ref).max() > 10000
    assert np.array_equal(out, ref), _diff(out, ref)
