"""Named parity scenarios (deterministic).  Each builder returns a Scenario.

The same cases feed: golden generation from the reference (tests/golden/
make_golden.py), the port-vs-golden CPU tests, and the CUDA-vs-oracle GPU tests.
Register layout: see scenarios.REGS (reference A2_crdesc order).
"""
import numpy as np

from scenarios import Scenario, fx

W, P, A, PH = 0, 1, 2, 3            # wtosc
VOL, PAN = 0, 1                      # panmix
CUT, Q, LP, BP, HP = 0, 1, 2, 3, 4   # filter12


def _rng(seed):
    return np.random.RandomState(seed)


def renderwave():
    """BASELINE config 1: one wtosc->panmix sine voice, mono, buffer 1024
    (mirrors test/renderwave.c:46-48 + PlayTestWave's envelope)."""
    s = Scenario(48000, 1, 1024, 4096)
    w = s.wave("sine")
    s.add_voice(["wtosc", "panmix"], [
        ("ramp", 0, W, w << 16), ("set", 0, P, fx(0.0)),
        ("ramp", 0, A, fx(1.0)), ("d", fx(10)),
        ("ramp", 0, A, fx(0.0)), ("d", fx(50)),
    ])
    return s


def osc_pan_ramps():
    s = Scenario(48000, 2, 64, 4800)
    w = s.wave("sine")
    s.add_voice(["wtosc", "panmix"], [
        ("ramp", 0, W, w << 16), ("set", 0, P, fx(0.0)),
        ("ramp", 0, A, fx(0.5)), ("d", fx(10)),
        ("ramp", 0, A, fx(0.1)), ("ramp", 1, PAN, fx(-0.5)), ("d", fx(33.3)),
        ("ramp", 1, PAN, fx(1.5)), ("ramp", 1, VOL, fx(0.7)), ("d", fx(20)),
        ("ramp", 0, P, fx(2.0)), ("d", fx(25)),
    ])
    return s


def all_waves():
    """Every builtin wave type at several pitches incl. mip switches, the
    muted range (> 11 octaves) and a phase write."""
    s = Scenario(48000, 2, 256, 3000)
    names = ["pulse2", "pulse10", "pulse25", "square", "saw", "triangle",
             "sine", "asine", "hsine", "qsine"]
    for i, n in enumerate(names):
        w = s.wave(n)
        p0 = -3.0 + 0.9 * i
        s.add_voice(["wtosc", "panmix"], [
            ("ramp", 0, W, w << 16), ("set", 0, P, fx(p0)),
            ("set", 0, A, fx(0.2)), ("set", 1, PAN, fx(-1 + 0.2 * i)),
            ("d", fx(7.5)),
            ("ramp", 0, P, fx(p0 + 4.0)), ("d", fx(20)),
            ("set", 0, PH, fx(0.25)), ("ramp", 0, P, fx(9.5)), ("d", fx(15)),
            ("ramp", 0, P, fx(-6.0)), ("d", fx(10)),
        ])
    return s


def filter_sweep():
    s = Scenario(48000, 2, 64, 6000)
    w = s.wave("saw")
    for i in range(6):
        p0 = -1.0 + 0.5 * i
        s.add_voice(["wtosc", "filter12", "panmix"], [
            ("ramp", 0, W, w << 16), ("set", 0, P, fx(p0)),
            ("set", 0, A, fx(0.3)),
            ("set", 1, CUT, fx(p0 + 1)), ("set", 1, Q, fx(2 + i)),
            ("set", 1, LP, fx(1.0 if i % 3 == 0 else 0.0)),
            ("set", 1, BP, fx(1.0 if i % 3 == 1 else 0.25)),
            ("set", 1, HP, fx(1.0 if i % 3 == 2 else 0.0)),
            ("set", 2, PAN, fx(-0.8 + 0.3 * i)),
            ("d", fx(5)),
            ("ramp", 1, CUT, fx(p0 + 5)), ("d", fx(40)),
            ("ramp", 1, CUT, fx(p0 - 1)), ("ramp", 1, Q, fx(0.3)), ("d", fx(30.7)),
            ("ramp", 1, CUT, fx(9.0)), ("d", fx(20)),
        ])
    return s


def additive8():
    """BASELINE config 3 shape: 8 x wtosc + panmix per voice, buffer 256."""
    s = Scenario(48000, 2, 256, 2560)
    w = s.wave("sine")
    r = _rng(3)
    for v in range(5):
        p0 = float(r.randint(-2 * 64, 2 * 64)) / 64
        steps = []
        for k in range(8):
            steps += [("ramp", k, W, w << 16),
                      ("set", k, P, fx(p0 + np.log2(k + 1))),
                      ("set", k, A, fx(0.1 / (k + 1)))]
        steps += [("set", 8, PAN, fx(float(r.randint(-64, 64)) / 64)),
                  ("d", fx(10))]
        for k in range(8):
            steps.append(("ramp", k, A, fx(0.02)))
        steps.append(("d", fx(30)))
        s.add_voice(["wtosc"] * 8 + ["panmix"], steps)
    return s


from audiality2_b200.workloads import FM_SETTINGS as _FM_SETTINGS  # noqa: E402


def fm_steps(kind, pitch, vel, pan):
    st = _FM_SETTINGS[kind]
    steps = [("ramp", 0, 2, fx(vel * st[0][0])), ("set", 0, 1, fx(pitch)),
             ("ramp", 0, 3, fx(st[0][1]))]
    for o in range(1, len(st)):
        p, a, fb = st[o]
        steps += [("set", 0, 2 + 3 * o, fx(a)), ("set", 0, 1 + 3 * o, fx(p)),
                  ("ramp", 0, 3 + 3 * o, fx(fb))]
    steps += [("set", 1, PAN, fx(pan)), ("d", fx(5))]
    return steps


def fm_all():
    s = Scenario(48000, 2, 64, 4000)
    for i, kind in enumerate(["fm1", "fm2", "fm3", "fm4", "fm3p", "fm4p",
                              "fm2r", "fm4r"]):
        steps = fm_steps(kind, -1.0 + 0.25 * i, 0.5, -0.7 + 0.2 * i)
        n = len(_FM_SETTINGS[kind])
        steps += [("ramp", 0, 2, fx(0.2)), ("ramp", 0, 3, fx(0.1)),
                  ("ramp", 0, 1, fx(0.5 + 0.25 * i))]
        if n > 1:
            steps += [("ramp", 0, 5, fx(0.1)), ("ramp", 0, 4, fx(2.0))]
        steps += [("d", fx(40)), ("set", 0, 0, fx(0.3)),
                  ("ramp", 0, 2, fx(0.0)), ("d", fx(20))]
        s.add_voice([kind, "panmix"], steps)
    return s


def waveshaper():
    s = Scenario(48000, 2, 64, 3000)
    w = s.wave("sine")
    for i in range(4):
        s.add_voice(["wtosc", "waveshaper", "panmix"], [
            ("ramp", 0, W, w << 16), ("set", 0, P, fx(-1.0 + i)),
            ("set", 0, A, fx(0.8)), ("set", 1, 0, fx(0.5 * i)),
            ("set", 2, PAN, fx(-0.5 + 0.3 * i)), ("d", fx(5)),
            ("ramp", 1, 0, fx(5.0 + 10 * i)), ("d", fx(30)),
            ("ramp", 1, 0, fx(0.0)), ("ramp", 0, A, fx(0.1)), ("d", fx(20)),
        ])
    return s


def mono_voice_and_chains():
    """{wtosc} (1 channel straight to the bus), {wtosc; wtosc; filter12},
    {fm2; waveshaper; panmix}."""
    s = Scenario(48000, 2, 64, 3000)
    sq = s.wave("square")
    tri = s.wave("triangle")
    s.add_voice(["wtosc"], [
        ("ramp", 0, W, sq << 16), ("set", 0, P, fx(0.5)),
        ("ramp", 0, A, fx(0.3)), ("d", fx(12)),
        ("ramp", 0, A, fx(0.0)), ("d", fx(40))])
    s.add_voice(["wtosc", "wtosc", "filter12"], [
        ("ramp", 0, W, sq << 16), ("ramp", 1, W, tri << 16),
        ("set", 0, P, fx(-1.0)), ("set", 1, P, fx(-0.99)),
        ("set", 0, A, fx(0.2)), ("set", 1, A, fx(0.2)),
        ("set", 2, CUT, fx(2.0)), ("set", 2, Q, fx(4)), ("d", fx(8)),
        ("ramp", 2, CUT, fx(-1.0)), ("d", fx(45))])
    s.add_voice(["fm2", "waveshaper", "panmix"], fm_steps("fm2", 0.0, 0.6, 0.3)
                [:-2] + [("set", 1, 0, fx(3.0)), ("set", 2, PAN, fx(0.3)),
                         ("d", fx(5)), ("ramp", 1, 0, fx(0.2)), ("d", fx(50))])
    return s


def groups():
    s = Scenario(48000, 2, 64, 4000)
    saw = s.wave("saw")
    g0 = s.add_group([("set", 0, VOL, fx(0.5)), ("set", 0, PAN, fx(-0.5)),
                      ("d", fx(10.3)), ("ramp", 0, VOL, fx(1.0)),
                      ("ramp", 0, PAN, fx(0.9)), ("d", fx(40))])
    g1 = s.add_group([("set", 0, VOL, fx(0.8)), ("d", fx(25)),
                      ("ramp", 0, PAN, fx(-1.4)), ("d", fx(30))])
    r = _rng(7)
    for i in range(12):
        s.add_voice(["wtosc", "panmix"], [
            ("ramp", 0, W, saw << 16),
            ("set", 0, P, fx(float(r.randint(-128, 128)) / 64)),
            ("set", 0, A, fx(0.1)),
            ("set", 1, PAN, fx(float(r.randint(-64, 64)) / 64)),
            ("d", fx(5 + i)), ("ramp", 0, A, fx(0.02)), ("d", fx(50)),
        ], group=[g0, g1, -1][i % 3])
    return s


def song_fbdelay():
    """Two song-level chains { inline 0 *; fbdelay * *; panmix * > } (the structure of every song of
    the reference, benchmark/k2trance.a2s:920-924) fed by plucked voices, with short delays (taps
    inside the current fragment: the serial path of the device fbdelay) and long ones (parallel
    path), a volume ramp on the song panmix and a voice outside the songs. Drop-in mode and the
    oracle port only: bank mode has no group-level fbdelay."""
    s = Scenario(48000, 2, 64, 9600)
    saw = s.wave("saw")
    sq = s.wave("square")
    # fbdelay, ldelay, rdelay in ms; drygain, fbgain, lgain, rgain
    g0 = s.add_group([("set", 0, VOL, fx(0.9)), ("d", fx(60)), ("ramp", 0, VOL, fx(0.4)),
                      ("ramp", 0, PAN, fx(0.5)), ("d", fx(80))],
                     fbdelay=[fx(31.7), fx(23.1), fx(41.9), fx(1.0), fx(0.35), fx(0.3), fx(0.3)])
    g1 = s.add_group([("set", 0, VOL, fx(0.7)), ("set", 0, PAN, fx(-0.3))],
                     fbdelay=[fx(0.9), fx(0.4), fx(1.1), fx(0.8), fx(0.5), fx(0.4), fx(0.45)])
    r = _rng(21)
    for i in range(10):
        s.add_voice(["wtosc", "panmix"], [
            ("ramp", 0, W, (saw if i & 1 else sq) << 16),
            ("set", 0, P, fx(float(r.randint(-128, 128)) / 64)),
            ("set", 0, A, fx(0.12)),
            ("set", 1, PAN, fx(float(r.randint(-64, 64)) / 64)),
            ("d", fx(3 + 2 * i)), ("ramp", 0, A, fx(0.0)), ("d", fx(25)),
            ("set", 0, A, fx(0.1)), ("ramp", 0, A, fx(0.0)), ("d", fx(40)),
        ], group=[g0, g1, g0, g1, -1][i % 5])
    return s


def rate44k_transposed():
    """44.1 kHz: ms delays land on fractional frames -> sub-sample starts."""
    s = Scenario(44100, 2, 100, 5000)
    s.transpose = fx(0.25)
    w = s.wave("triangle")
    saw = s.wave("saw")
    s.add_voice(["wtosc", "panmix"], [
        ("ramp", 0, W, w << 16), ("set", 0, P, fx(0.3)),
        ("ramp", 0, A, fx(0.6)), ("d", fx(3.7)),
        ("ramp", 0, P, fx(1.3)), ("d", fx(11.1)),
        ("ramp", 0, A, fx(0.0)), ("ramp", 1, PAN, fx(0.4)), ("d", fx(41.3)),
    ])
    s.add_voice(["wtosc", "filter12", "panmix"], [
        ("ramp", 0, W, saw << 16), ("set", 0, P, fx(-0.7)),
        ("set", 0, A, fx(0.4)), ("set", 1, CUT, fx(1.0)), ("set", 1, Q, fx(3)),
        ("d", fx(2.3)), ("ramp", 1, CUT, fx(4.0)), ("d", fx(33.3)),
        ("ramp", 1, CUT, fx(0.0)), ("d", fx(21.7)),
    ])
    s.add_voice(["fm3", "panmix"],
                fm_steps("fm3", -0.5, 0.5, 0.2)[:-1] + [("d", fx(7.7)),
                ("ramp", 0, 2, fx(0.0)), ("d", fx(60.1))])
    return s


def noise():
    s = Scenario(48000, 2, 64, 3000, noiseseed=4711)
    n = s.wave("noise")
    sine = s.wave("sine")
    for i in range(3):
        s.add_voice(["wtosc", "panmix"], [
            ("ramp", 0, W, n << 16), ("set", 0, P, fx(1.0 + 2.5 * i)),
            ("set", 0, A, fx(0.3)), ("set", 1, PAN, fx(-0.5 + 0.5 * i)),
            ("d", fx(10)), ("ramp", 0, P, fx(6.0 - 2 * i)), ("d", fx(30)),
            ("ramp", 0, A, fx(0.0)), ("d", fx(15)),
        ])
    s.add_voice(["wtosc", "panmix"], [
        ("ramp", 0, W, sine << 16), ("set", 0, P, fx(0.0)),
        ("set", 0, A, fx(0.2)), ("d", fx(50))])
    return s


def bank(nvoices=256, kinds=("wtosc", "filter12", "panmix"), wave="saw",
         seed=324357, frames=1280, buffer=64):
    """BASELINE config 2 shape at a size the oracle finishes in seconds:
    static voices, pitch uniform in [-2, 2) octaves, pan in [-1, 1)."""
    s = Scenario(48000, 2, buffer, frames)
    w = s.wave(wave)
    r = _rng(seed)
    nosc = sum(1 for k in kinds if k == "wtosc")
    for v in range(nvoices):
        p0 = int(r.randint(-2 * 65536, 2 * 65536))
        pan = int(r.randint(-65536, 65536))
        steps = []
        for k in range(nosc):
            steps += [("ramp", k, W, w << 16),
                      ("set", k, P, p0 + fx(np.log2(k + 1))),
                      ("set", k, A, fx(0.0002 * 8 / (k + 1)))]
        ui = nosc
        if "filter12" in kinds:
            steps += [("set", ui, CUT, p0 + 65536), ("set", ui, Q, fx(2))]
            ui += 1
        steps += [("set", ui, PAN, pan), ("d", fx(1000))]
        s.add_voice(list(kinds), steps)
    return s


def bench_bank(nvoices=4096, steps=4, step_ms=20, seed=324357):
    """The bench.py workload (BASELINE config 2 + one control write per voice
    per step): static pitch/cutoff/pan, amplitude re-targeted every step_ms
    and ramped across the step, 48 kHz, 64-frame blocks."""
    from audiality2_b200.workloads import cfg2_bank
    b = cfg2_bank(nvoices, seed)
    s = Scenario(48000, 2, 64, steps * step_ms * 48)
    w = s.wave(b["wave"])
    a0 = b["amp"]
    for v in range(nvoices):
        s.add_voice(list(b["kinds"]), [
            ("ramp", 0, W, w << 16), ("set", 0, P, int(b["pitch"][v])),
            ("set", 0, A, a0),
            ("set", 1, CUT, int(b["cutoff"][v])), ("set", 1, Q, b["q"]),
            ("set", 2, PAN, int(b["pan"][v])),
            ("loop", (steps + 1) // 2, [
                ("ramp", 0, A, a0 // 2), ("d", fx(step_ms)),
                ("ramp", 0, A, a0), ("d", fx(step_ms))]),
        ])
    return s


def fm_bank(nvoices=64, frames=1280):
    """BASELINE config 4 shape: voices cycling fm3/fm3p/fm2r/fm4r."""
    s = Scenario(48000, 2, 64, frames)
    r = _rng(11)
    kinds = ["fm3", "fm3p", "fm2r", "fm4r"]
    for v in range(nvoices):
        k = kinds[v % 4]
        p0 = float(r.randint(-2 * 64, 2 * 64)) / 64
        pan = float(r.randint(-64, 64)) / 64
        s.add_voice([k, "panmix"], fm_steps(k, p0, 0.05, pan)[:-1] +
                    [("d", fx(1000))])
    return s



def _sampled(wtype, period, flags, length, seed):
    """An uploaded (sampled) wave played from -3 to +8.5 octaves: below and above A2_MAXPHINC
    samples per frame (plain loop, per-sample wrapped loop / end check, muted range), with a pitch
    ramp and a phase write. The reference's harness uploads the same data (a2render -U)."""
    s = Scenario(48000, 2, 64, 2400)
    w = s.upload(wtype, period, flags, length, seed)
    for i in range(12):
        p0 = -3.0 + 1.05 * i
        s.add_voice(["wtosc", "panmix"], [
            ("ramp", 0, W, w << 16), ("set", 0, P, fx(p0)),
            ("set", 0, A, fx(0.15)), ("set", 1, PAN, fx(-0.9 + 0.15 * i)),
            ("d", fx(9.25)),
            ("ramp", 0, P, fx(p0 + 1.5)), ("ramp", 0, A, fx(0.05)), ("d", fx(21)),
            ("set", 0, PH, fx(0.5)), ("d", fx(12)),
        ])
    return s


def sampled_loop():
    """Looped non-mipmapped wave (A2_WWAVE | A2_LOOPED), 3001 samples, period 500."""
    return _sampled(2, 500, 0x100, 3001, 7)


def sampled_oneshot():
    """Non-looped non-mipmapped wave: voices run out of samples mid-window."""
    return _sampled(2, 500, 0, 3001, 7)


def sampled_mip():
    """Uploaded mipmapped wave (A2_WMIPWAVE | A2_LOOPED): mip levels rendered by the host."""
    return _sampled(3, 64, 0x100, 5000, 9)


CASES = {
    "renderwave": renderwave,
    "osc_pan_ramps": osc_pan_ramps,
    "all_waves": all_waves,
    "filter_sweep": filter_sweep,
    "additive8": additive8,
    "fm_all": fm_all,
    "waveshaper": waveshaper,
    "mono_voice_and_chains": mono_voice_and_chains,
    "groups": groups,
    "song_fbdelay": song_fbdelay,
    "rate44k_transposed": rate44k_transposed,
    "noise": noise,
    "sampled_loop": sampled_loop,
    "sampled_oneshot": sampled_oneshot,
    "sampled_mip": sampled_mip,
    "bank256": bank,
    "fm_bank64": fm_bank,
    "bench_bank200": lambda: bench_bank(200, steps=3),
}


# Hand-written scripts rendered as whole songs: name -> (path, program, frames, rate, buffer)
SCRIPTS = {
    "hybrid_song": ("data/hybrid_song.a2s", "Song", 30000, 48000, 64),
    "hybrid_song_44k_b256": ("data/hybrid_song.a2s", "Song", 20000, 44100, 256),
    # replaced units outside fused leaf voices (generic per-unit device ops)
    "generic_chains": ("data/generic_chains.a2s", "Song", 40000, 48000, 64),
    "generic_chains_44k_b200": ("data/generic_chains.a2s", "Song", 30000, 44100, 200),
    # limiter / dcblock / dc as device units
    "bus_effects": ("data/bus_effects.a2s", "Song", 20000, 48000, 64),
    "bus_effects_44k_b96": ("data/bus_effects.a2s", "Song", 16000, 44100, 96),
}
