"""Synthetic voice banks of BASELINE.json's configs (parameters only; all values
16:16 fixed point).  Shared by bench.py and the parity tests so that what is
measured is exactly what is checked."""
import numpy as np

FX_ONE = 65536


def fx(x):
    return int(np.floor(x * 65536.0 + 0.5))


def cfg2_bank(nvoices=4096, seed=324357):
    """BASELINE config 2: voices wtosc -> filter12 -> panmix on the shared
    2048-point saw; pitch uniform in [-2, 2) octaves, cutoff one octave above
    the pitch, q 2, pan uniform in [-1, 1), amplitude 0.0002."""
    r = np.random.RandomState(seed)
    p0 = np.empty(nvoices, dtype=np.int32)
    pan = np.empty(nvoices, dtype=np.int32)
    for v in range(nvoices):            # same draw order as tests/cases.py
        p0[v] = r.randint(-2 * 65536, 2 * 65536)
        pan[v] = r.randint(-65536, 65536)
    return {
        "kinds": ("wtosc", "filter12", "panmix"),
        "wave": "saw",
        "pitch": p0,
        "cutoff": p0 + FX_ONE,
        "q": fx(2),
        "pan": pan,
        "amp": fx(0.0002),
    }


# (a, fb) of operator 0, then (p, a, fb) per further operator: the shapes of the reference's
# fmtest4 instruments (benchmark/fmtest4.a2s:11-79)
FM_SETTINGS = {
    "fm1": [(1.0, 0.5)],
    "fm2": [(1.0, 0.4), (1.0, 0.8, 0.3)],
    "fm3": [(1.0, 0.7), (0.99, 0.5, 0.2), (1.01, 1.0, 0.2)],
    "fm4": [(1.0, 0.3), (1.0, 0.6, 0.2), (2.0, 0.5, 0.1), (3.01, 0.4, 0.3)],
    "fm3p": [(1.0, 0.7), (0.99, 0.5, 0.2), (1.01, 1.0, 0.2)],
    "fm4p": [(1.0, 0.5), (1.0, 0.5, 0.2), (1.98, 0.7, 0.5), (3.02, 0.5, 0.3)],
    "fm2r": [(1.0, 0.9), (1.01, 1.0, 0.8)],
    "fm4r": [(1.0, 0.6), (1.0, 1.0, 0.7), (1.98, 0.7, 0.5), (3.02, 0.5, 0.3)],
}
CFG4_KINDS = ("fm3", "fm3p", "fm2r", "fm4r")


def setup_cfg2(e, nvoices=4096, seed=324357, ramp_frames=960):
    """One cfg2 bank on engine `e` (audiality2_b200.engine.Engine). Returns (bank id, params)."""
    from .chains import autowire
    b = cfg2_bank(nvoices, seed)
    w = e.builtin_wave(b["wave"])
    bank = e.new_bank(autowire(list(b["kinds"])), nvoices)
    e.write_all(bank, 0, 0, [w << 16], dur=ramp_frames << 8)
    e.write_all(bank, 0, 1, b["pitch"])
    e.write_all(bank, 0, 2, [b["amp"]])
    e.write_all(bank, 1, 0, b["cutoff"])
    e.write_all(bank, 1, 1, [b["q"]])
    e.write_all(bank, 2, 1, b["pan"])
    return bank, b


def setup_cfg3(e, nvoices=65536, nosc=8, seed=3, writer=None):
    """BASELINE config 3: additive voices, `nosc` x wtosc (sine, partial k at p + log2(k), a ~ 1/k)
    + panmix. `writer(unit, reg, values, dur)` lets a test mirror every write into the oracle."""
    from .chains import autowire
    r = np.random.RandomState(seed)
    pitch = r.randint(-2 * 65536, 2 * 65536, size=nvoices).astype(np.int32)
    pan = r.randint(-65536, 65536, size=nvoices).astype(np.int32)
    chain = autowire(["wtosc"] * nosc + ["panmix"])
    w = e.builtin_wave("sine")
    bank = e.new_bank(chain, nvoices)

    def wr(unit, reg, values, dur=0, wave=False):
        e.write_all(bank, unit, reg, values, dur=dur)
        if writer:
            writer(unit, reg, values, dur, wave)

    for k in range(nosc):
        wr(k, 1, pitch + fx(np.log2(k + 1)))
        wr(k, 2, [fx(0.00002 / (k + 1))])
        wr(k, 0, [w << 16], wave=True)
    wr(nosc, 1, pan)
    return [bank], chain


def setup_cfg4(e, nvoices=32768, seed=11, first_voice=0, total=None, writer=None):
    """BASELINE config 4 (one GPU's shard): voices cycling the four fmtest4 instruments
    {fmX; panmix}, X in fm3 / fm3p / fm2r / fm4r, register settings held static, one ramping
    feedback per modulator. Voices [first_voice, first_voice + nvoices) of a bank of `total`
    voices whose per-voice pitch / pan come from one seeded stream (so shards of one bank agree
    with the whole). Returns the four bank ids."""
    from .chains import autowire
    total = total or (first_voice + nvoices)
    r = np.random.RandomState(seed)
    pitch_all = r.randint(-2 * 65536, 2 * 65536, size=total).astype(np.int32)
    pan_all = r.randint(-65536, 65536, size=total).astype(np.int32)
    banks = []
    for ki, kind in enumerate(CFG4_KINDS):
        idx = np.arange(first_voice, first_voice + nvoices)
        idx = idx[idx % 4 == ki]
        n = len(idx)
        chain = autowire([kind, "panmix"])
        bank = e.new_bank(chain, n)
        st = FM_SETTINGS[kind]

        def wr(unit, reg, values, dur=0):
            e.write_all(bank, unit, reg, values, dur=dur)
            if writer:
                writer(ki, chain, n, unit, reg, values, dur)

        wr(0, 1, pitch_all[idx])
        wr(0, 2, [fx(0.0005 * st[0][0])])
        wr(0, 3, [fx(st[0][1])])
        for op in range(1, len(st)):
            p, a, fb = st[op]
            wr(0, 1 + 3 * op, [fx(p)])
            wr(0, 2 + 3 * op, [fx(a)])
            wr(0, 3 + 3 * op, [fx(fb)], dur=100 << 8)
        wr(1, 1, pan_all[idx])
        banks.append(bank)
    return banks


def setup_gather(e, nvoices=131072, nwaves=12, samples_per_frame=64, length=(1 << 24) - 256, seed=1):
    """The HBM-bound wtosc gather (SURVEY.md 8(d) "honest caveat"): `nwaves` looped, non-mipmapped
    SAMPLED waves of 16 M samples each (the reference's limit is 2^24 - 133 frames, wtosc.c:55;
    together 3 x the 126 MB L2) played by {wtosc; panmix} voices at `samples_per_frame` wave
    samples per output frame from random start phases. At 64 samples per frame the two Hermite
    taps of one output sample (a2_Hermite at ph and ph + dph/2, wtosc.c:226-228) are 64 bytes
    apart and consecutive frames 128 bytes: every tap touches its own 32-byte sector, so the
    algorithmic HBM traffic is 2 sectors = 64 B per voice-sample."""
    from .chains import autowire
    rng = np.random.RandomState(seed)
    base = rng.randint(-20000, 20000, size=length).astype(np.int16)
    dphase0 = 261.626 / e.samplerate * (1 << 24)         # a2_P2I at pitch 0 incl. basepitch
    period = int(round(samples_per_frame * (1 << 24) / dphase0))
    waves = [e.upload_wave(2, period, 0x100, np.roll(base, 7919 * w)) for w in range(nwaves)]
    bank = e.new_bank(autowire(["wtosc", "panmix"]), nvoices)
    e.write_all(bank, 0, 0, (np.array([waves[v % nwaves] for v in range(nvoices)], dtype=np.int64) << 16)
                .astype(np.int32))
    e.write_all(bank, 0, 1, [0])
    e.write_all(bank, 0, 2, [65])
    periods = length // period
    e.write_all(bank, 0, 3, (rng.randint(0, periods - 2, size=nvoices).astype(np.int64) << 16).astype(np.int32))
    e.write_all(bank, 1, 1, rng.randint(-65536, 65536, size=nvoices).astype(np.int32))
    return [bank], {"bytes_per_voice_sample": 64, "wave_bytes": nwaves * length * 2, "period": period}
