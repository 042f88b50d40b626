"""Arbitrary voice structures in bank mode: render_generic (one thread per voice, unit by unit over
each segment, csrc/a2cu_bus.cuh) against the oracle port on randomly generated structs.

The reference's compiler accepts any `struct { ... }` whose units agree on 0-2 scratch channels
(src/compiler.c:2991-3188); the engine has fused kernels for the structures that carry the load and
runs everything else generically. 64 random chains: 1-4 generators (wtosc / fm*), up to four
processors (filter12, waveshaper in 1 or 2 channels, panmix 1->1 / 1->2 / 2->1 / 2->2, adding or
replacing, further generators in between), a wired-out panmix / filter12 / waveshaper at the end;
random register settings at time 0, ramps and jumps mid-render, several transposed voices per chain.
Bit-exact int32 against the port (pinned to the reference, tests/test_oracle.py)."""
import numpy as np
import pytest

from audiality2_b200.chains import KIND_CODE, REGS
from oracle import a2oracle as ao
from scenarios import fx

pytestmark = pytest.mark.gpu
NAME = {v: k for k, v in KIND_CODE.items()}
GENS = ["wtosc", "wtosc", "wtosc", "fm1", "fm2", "fm3", "fm3p", "fm2r", "fm4", "fm4p", "fm4r"]


def random_chain(r):
    units = []
    for g in range(r.randint(1, 4)):
        units.append([KIND_CODE[GENS[r.randint(len(GENS))]], 0, 1, 1 if g else 0, 0])
    ch = 1
    for _ in range(r.randint(0, 5)):
        k = ["filter12", "waveshaper", "pm12", "pm11", "pm21", "pm22", "gen"][r.randint(7)]
        add = 1 if r.randint(5) == 0 else 0
        if k == "pm12" and ch == 1:
            units.append([2, 1, 2, 0, 0]); ch = 2
        elif k == "pm11" and ch == 1:
            units.append([2, 1, 1, add, 0])
        elif k == "pm21" and ch == 2:
            units.append([2, 2, 1, 0, 0]); ch = 1
        elif k == "pm22" and ch == 2:
            units.append([2, 2, 2, add, 0])
        elif k in ("filter12", "waveshaper"):
            units.append([KIND_CODE[k], ch, ch, add, 0])
        elif k == "gen":
            units.append([KIND_CODE[GENS[r.randint(len(GENS))]], 0, 1, 1, 0])
    final = ["panmix", "panmix", "filter12", "waveshaper"][r.randint(4)]
    if final == "panmix":
        units.append([2, ch, r.randint(1, 3), 1, 1])
    else:
        units.append([KIND_CODE[final], ch, ch, 1, 1])
    return [tuple(u) for u in units]


def settings(r, kind, waves, t):
    """Random (reg, value, dur) writes for one unit at one moment."""
    name = NAME[kind]
    out = []
    d = 0 if t == 0 else [0, 0, 300 << 8, 700 << 8][r.randint(4)]
    if name == "wtosc":
        if t == 0:
            out.append((0, waves[r.randint(len(waves))] << 16, 0))
        out += [(1, fx(r.uniform(-2, 2.5)), d), (2, fx(r.uniform(0.02, 0.3)), d)]
        if r.randint(3) == 0:
            out.append((3, fx(r.uniform(0, 1)), 0))
    elif name == "panmix":
        out += [(0, fx(r.uniform(0.3, 1.0)), d), (1, fx(r.uniform(-1.2, 1.2)), d)]
    elif name == "filter12":
        out += [(0, fx(r.uniform(-1, 3)), d), (1, fx(r.uniform(0.8, 4)), d)]
        if t == 0:
            out += [(2, fx(r.uniform(0, 1)), 0), (3, fx(r.uniform(0, 1)), 0), (4, fx(r.uniform(0, 0.5)), 0)]
    elif name == "waveshaper":
        out.append((0, fx(r.uniform(0, 2)), d))
    else:       # fm*: phase, then (p, a, fb) per operator
        nreg = len(REGS[name])
        out += [(1, fx(r.uniform(-1.5, 1.5)), d), (2, fx(r.uniform(0.05, 0.4)), d), (3, fx(r.uniform(0, 0.5)), d)]
        for op in range(1, (nreg - 1) // 3):
            out += [(1 + 3 * op, fx(r.uniform(0.5, 3)), 0 if t == 0 else d), (2 + 3 * op, fx(r.uniform(0, 1)), d),
                    (3 + 3 * op, fx(r.uniform(0, 0.4)), d)]
    return out


@pytest.mark.parametrize("seed", range(64))
def test_random_structure_matches_port(seed):
    from audiality2_b200 import engine as eng
    r = np.random.RandomState(1000 + seed)
    chain = random_chain(r)
    V, frames, buffer = 12, 1600, [64, 64, 100, 256][seed % 4]
    e = eng.Engine(48000, 2)
    o = ao.Oracle(48000, 2)
    try:
        waves_e = [e.builtin_wave(n) for n in ("sine", "saw", "triangle", "pulse25")]
        waves_o = [o.builtin_wave(n) for n in ("sine", "saw", "triangle", "pulse25")]
        assert waves_e == waves_o
        transposes = [fx(r.uniform(-1, 1)) for _ in range(V)]
        bank = e.new_bank(chain, V, transpose=transposes)
        for v in range(V):
            o.new_voice(chain, transpose=transposes[v])
        events = []
        times = [0, (480 << 8) + 37, (777 << 8) + 200, 1203 << 8]
        for t in times:
            for v in range(V):
                for u, spec in enumerate(chain):
                    if t and r.randint(3):
                        continue
                    for reg, value, dur in settings(r, spec[0], waves_e, t):
                        e.write(bank, v, u, reg, value, t, dur)
                        events.append((t, ao.EV_WRITE, v, u, reg, value, dur))
                if t:
                    e.wake(bank, v, t)
                    events.append((t, ao.EV_WAKE, v, 0, 0, 0, 0))
        t = 1000000        # the root's own wake-ups (engine: root_wake_period)
        while t < frames << 8:
            events.append((t, ao.EV_ROOTWRITE, 0, 0, -1, 0, 0))
            t += 1000000
        order = sorted(range(len(events)), key=lambda i: events[i][0])
        arr = np.zeros(len(events), dtype=ao.EVENT_DTYPE)
        for j, i in enumerate(order):
            arr[j] = events[i]
        out = e.run(frames, buffer)
        ref = o.render(arr, frames, buffer)
        kname = e.bank_kernel_name(bank)
    finally:
        e.close()
        o.close()
    assert np.abs(ref).max() > 1000, "silent chain %r" % (chain,)
    if not np.array_equal(out, ref):
        bad = np.nonzero((out != ref).any(axis=1))[0]
        raise AssertionError("chain %r (%s): first diff at frame %d, %d frames differ, max abs %d" % (
            chain, kname, bad[0], len(bad), np.abs(out.astype(np.int64) - ref).max()))


def test_generic_kernel_is_what_ran():
    """At least the odd shapes must really take render_generic (no fused kernel exists for them)."""
    from audiality2_b200 import engine as eng
    e = eng.Engine(48000, 2)
    try:
        chain = [(1, 0, 1, 0, 0), (2, 1, 2, 0, 0), (3, 2, 2, 0, 0), (4, 2, 2, 1, 0), (2, 2, 2, 1, 1)]
        assert eng.Engine.chain_supported(chain)
        bank = e.new_bank(chain, 4)
        assert e.bank_kernel_name(bank) == "generic"
        # fbdelay is a cooperative bus unit, not a per-voice one
        assert not eng.Engine.chain_supported([(1, 0, 1, 0, 0), (5, 1, 1, 0, 0), (2, 1, 2, 1, 1)])
    finally:
        e.close()
