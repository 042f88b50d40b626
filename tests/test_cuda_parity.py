"""GPU parity tests: the CUDA engine, called through its C ABI, against the
plain-C oracle port on the same inputs and against the golden vectors that
came from the reference build.  Bar: bit-exact int32."""
import os

import numpy as np
import pytest

from cases import CASES, bank, fm_bank
from scenarios import run_cuda, run_oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_outputs.npz")

# the noise oscillator shares one LCG across voices in tree-walk order
# (wtosc.c:135-144); it has its own test below
# song_fbdelay: group-level fbdelay exists in drop-in mode only (tests/test_dropin.py)
PARITY_CASES = [n for n in sorted(CASES) if n not in ("noise", "song_fbdelay")]


def _diff(a, b):
    bad = np.nonzero((a != b).any(axis=1))[0]
    return "first diff at frame %d (%d frames differ, max abs %d)" % (
        bad[0], len(bad), np.abs(a.astype(np.int64) - b).max())


@pytest.mark.parametrize("name", PARITY_CASES)
def test_cuda_matches_golden_and_oracle(name):
    scn = CASES[name]()
    out = run_cuda(scn)
    ref = np.load(GOLDEN)[name]
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), _diff(out, ref)
    assert np.array_equal(out, run_oracle(scn))


@pytest.mark.parametrize("name", ["osc_pan_ramps", "filter_sweep", "fm_all", "groups"])
def test_windowing_does_not_change_output(name):
    """Rendering in several a2cu_run calls equals one call (state carried in HBM)."""
    scn = CASES[name]()
    one = run_cuda(scn)
    many = run_cuda(scn, window=scn.buffer * 5)
    assert np.array_equal(one, many), _diff(one, many)


@pytest.mark.parametrize("name", ["filter_sweep", "groups", "noise"])
def test_pipelined_submit_collect_equals_run(name):
    """a2cu_submit / a2cu_collect with three windows in flight (staging ring,
    per-slot pinned results) equals one synchronous a2cu_run."""
    scn = CASES[name]()
    one = run_cuda(scn)
    piped = run_cuda(scn, window=scn.buffer * 3, pipelined=True)
    assert np.array_equal(one, piped), _diff(one, piped)


def test_bank_4096_full_size():
    """BASELINE config 2 at full size (4096 voices) vs the oracle port."""
    scn = bank(4096, frames=640)
    out = run_cuda(scn)
    ref = run_oracle(scn)
    assert np.array_equal(out, ref), _diff(out, ref)


def test_additive_bank():
    scn = bank(512, kinds=("wtosc",) * 8 + ("panmix",), wave="sine", frames=512, buffer=256)
    out = run_cuda(scn)
    ref = run_oracle(scn)
    assert np.array_equal(out, ref), _diff(out, ref)


def test_fm_bank_1024():
    scn = fm_bank(1024, frames=512)
    out = run_cuda(scn)
    ref = run_oracle(scn)
    assert np.array_equal(out, ref), _diff(out, ref)


@pytest.mark.parametrize("name", ["bank256", "bench_bank200", "additive8", "filter_sweep",
                                  "osc_pan_ramps", "groups", "all_waves"])
def test_split_kernel_equals_thread_per_voice_kernel(name):
    """render_split (warp-specialised, closed-form oscillator/panmix stages)
    and render_bank (one thread per voice) on the same state layout."""
    scn = CASES[name]()
    st = {}
    a = run_cuda(scn, split=True, stats=st)
    b = run_cuda(scn, split=False)
    assert np.array_equal(a, b), _diff(a, b)
    if name in ("bank256", "bench_bank200"):
        assert st["split_launches"] > 0, "split kernel was expected to be eligible"


def test_linearity_of_bus():
    """Size-independent property: the bus is an integer sum, so rendering two
    disjoint halves of a bank separately and adding equals rendering it whole
    (holds exactly while the root panmix is the identity)."""
    full = bank(1024, frames=256)
    a = bank(1024, frames=256)
    b = bank(1024, frames=256)
    a.voices = a.voices[:512]
    b.voices = b.voices[512:]
    whole = run_cuda(full)
    parts = run_cuda(a) + run_cuda(b)
    assert np.array_equal(whole, parts)


def test_noise_oscillator():
    """Shared-LCG noise (wtosc.c:129-152): seeds come from the host planner in
    tree-walk order; three noise voices + pitch ramps + a plain voice."""
    scn = CASES["noise"]()
    out = run_cuda(scn)
    ref = np.load(GOLDEN)["noise"]
    assert np.array_equal(out, ref), _diff(out, ref)


def test_empty_engine_renders_silence():
    from audiality2_b200 import engine as eng
    e = eng.Engine()
    out = e.run(200, 96)
    assert out.shape == (200, 2) and not out.any()
    e.close()


def test_paused_bank_keeps_state_and_pending_writes():
    """a2cu_bank_enable: a paused bank is not rendered, keeps its voices' state
    and applies its pending writes when it runs again (bench.py serves many
    banks round-robin this way)."""
    from audiality2_b200 import engine as eng
    from audiality2_b200.workloads import cfg2_bank
    from scenarios import autowire
    W = 640

    def setup(e, seed):
        b = cfg2_bank(64, seed=seed)
        bank = e.new_bank(autowire(list(b["kinds"])), 64)
        e.write_all(bank, 0, 0, [e.builtin_wave(b["wave"]) << 16])
        e.write_all(bank, 0, 1, b["pitch"])
        e.write_all(bank, 0, 2, [b["amp"] * 50], dur=W << 8)
        e.write_all(bank, 1, 0, b["cutoff"])
        e.write_all(bank, 1, 1, [b["q"]])
        e.write_all(bank, 2, 1, b["pan"])
        return bank

    e1 = eng.Engine(48000, 2)
    setup(e1, 1)
    ref = [e1.run(W, 64) for _ in range(3)]
    e1.close()

    e2 = eng.Engine(48000, 2)
    x = setup(e2, 1)
    y = setup(e2, 2)
    e2.bank_enable(y, False)
    a = e2.run(W, 64)                   # x alone
    e2.bank_enable(x, False)
    e2.bank_enable(y, True)
    other = e2.run(W, 64)               # y's first window: its writes from time 0 apply now
    e2.bank_enable(y, False)
    e2.bank_enable(x, True)
    b2 = e2.run(W, 64)                  # x resumes where it stopped
    c2 = e2.run(W, 64)
    e2.close()
    assert np.abs(ref[0]).max() > 1000 and np.abs(other).max() > 1000
    assert np.array_equal(a, ref[0]), _diff(a, ref[0])
    assert np.array_equal(b2, ref[1]), _diff(b2, ref[1])
    assert np.array_equal(c2, ref[2]), _diff(c2, ref[2])
    assert not np.array_equal(other, ref[0])


@pytest.mark.parametrize("name", ["bank256", "groups", "renderwave"])
def test_output_formats_are_the_driver_edge_conversions(name):
    """a2cu_set_output_format: float32 = v * 2^-23 (drivers/sdldrv.c:55-65) and int16 = v >> 8
    (waves.c:174-176), produced inside the root stage, equal the conversions applied to the int32
    render - through a2cu_run, a windowed a2cu_run and the pipelined submit / collect path."""
    from audiality2_b200 import engine as eng
    from audiality2_b200.chains import autowire
    scn = CASES[name]()
    ref = run_cuda(scn)

    def render(fmt, pipelined):
        e = eng.Engine(scn.samplerate, scn.channels)
        try:
            e.set_output_format(fmt)
            for w in scn.waves:
                e.builtin_wave(w)
            for _ in range(scn.ngroups):
                e.new_group()
            banks = {}
            where = []
            for v in scn.voices:
                key = tuple(v.kinds)
                banks.setdefault(key, []).append(v)
                where.append((key, len(banks[key]) - 1))
            ids = {k: e.new_bank(autowire(list(k)), len(vs), transpose=[v.transpose for v in vs],
                                 group=[v.group for v in vs]) for k, vs in banks.items()}
            from oracle import a2oracle as ao
            for ev in scn.events():
                t, kind, tgt = int(ev["time"]), int(ev["kind"]), int(ev["voice"])
                if kind == ao.EV_WRITE:
                    k, slot = where[tgt]
                    e.write(ids[k], slot, int(ev["unit"]), int(ev["reg"]), int(ev["value"]), t, int(ev["dur"]))
                elif kind == ao.EV_WAKE:
                    k, slot = where[tgt]
                    e.wake(ids[k], slot, t)
                elif kind == ao.EV_GROUPWRITE:
                    e.group_write(tgt, int(ev["reg"]), int(ev["value"]), t, int(ev["dur"]))
            if not pipelined:
                return e.run(scn.frames, scn.buffer)
            parts, tickets, done = [], [], 0
            while done < scn.frames:
                n = min(scn.buffer * 3, scn.frames - done)
                tickets.append(e.submit(n, scn.buffer))
                done += n
                if len(tickets) > 2:
                    parts.append(e.collect(tickets.pop(0)))
            while tickets:
                parts.append(e.collect(tickets.pop(0)))
            return np.concatenate(parts, axis=0)
        finally:
            e.close()

    f32 = ref.astype(np.float32) * np.float32(1.0 / 8388608.0)
    i16 = (ref >> 8).astype(np.int16)
    for pipelined in (False, True):
        a = render("f32", pipelined)
        assert a.dtype == np.float32 and np.array_equal(a, f32)
        b = render("i16", pipelined)
        assert b.dtype == np.int16 and np.array_equal(b, i16)


@pytest.mark.parametrize("wtype,flags,length", [(3, 0x100, 2048), (3, 0x100, 5000), (3, 0, 3001), (2, 0x100, 3001),
                                                (2, 0, 777), (3, 0x100, 1)])
def test_device_wave_preparation_equals_port(wtype, flags, length):
    """Pads and mip levels are built by kernels in the device pool (csrc/a2cu_waves.cuh,
    waves.c:90-151); every level, pads included, must equal what the port prepares on the CPU."""
    from audiality2_b200 import engine as eng
    from oracle import a2oracle as ao
    r = np.random.RandomState(length)
    data = r.randint(-32768, 32768, size=length).astype(np.int16)
    e = eng.Engine(48000, 2)
    o = ao.Oracle(48000, 2)
    try:
        we = e.upload_wave(wtype, 64, flags, data)
        wo = o.upload_wave(wtype, 64, flags, data)
        for lvl in range(10 if wtype == 3 else 1):
            a, na = e.wave_data(we, lvl)
            b, nb = o.wave_data(wo, lvl)
            assert na == nb
            assert np.array_equal(a[:1 + na + 132], b[:1 + nb + 132]), "level %d differs" % lvl
        for name in ("saw", "pulse25", "sine", "triangle"):
            be, bo = e.builtin_wave(name), o.builtin_wave(name)
            for lvl in range(10):
                a, na = e.wave_data(be, lvl)
                b, nb = o.wave_data(bo, lvl)
                assert na == nb and np.array_equal(a[:1 + na + 132], b[:1 + nb + 132]), (name, lvl)
    finally:
        e.close()
        o.close()


def test_one_second_window_in_one_call():
    """a2cu_run with a 48 000-frame window: more root wake-ups than one launch takes and far more
    frames than render_split's per-launch limit, so the engine renders sub-windows, each into its own
    part of the output block (int32 and the half-size int16 format) - must equal the port."""
    from audiality2_b200 import engine as eng
    scn = bank(96, frames=48000)
    ref = run_oracle(scn)
    one = run_cuda(scn)                      # one a2cu_run(48000)
    assert np.array_equal(one, ref), _diff(one, ref)
    piped = run_cuda(scn, window=16000, pipelined=True)
    assert np.array_equal(piped, ref), _diff(piped, ref)
    assert np.abs(ref[40000:]).max() > 1000
