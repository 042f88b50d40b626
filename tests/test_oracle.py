"""CPU tests: the plain-C port (oracle/a2_oracle.c) against the golden vectors
generated from the reference, and - when the reference build is present
(oracle/_ref, dev container only) - against the reference itself, live."""
import ctypes as C
import os

import numpy as np
import pytest

from cases import CASES
from scenarios import run_oracle, run_ref
from oracle import a2oracle as ao

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_outputs.npz")
REF_LIB = os.path.join(ao.REF_DIR, "libaudiality2.so")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_golden(name, golden):
    out = run_oracle(CASES[name]())
    ref = golden[name]
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), "port differs from the reference's output (bit-exact bar)"
    assert np.abs(ref).max() > 1000, "degenerate golden (silence)"


@pytest.mark.skipif(not ao.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_live(name, tmp_path):
    scn = CASES[name]()
    ref = run_ref(scn, str(tmp_path / (name + ".a2s")))
    out = run_oracle(scn)
    assert np.array_equal(out, ref)


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built")
def test_p2i_matches_reference():
    R = C.CDLL(REF_LIB)
    R.a2_pitch_open()
    R.a2_P2I.restype = C.c_uint
    L = ao.lib()
    pitches = list(range(-14 * 65536, 14 * 65536, 1013))
    pitches += list(range(-7 * 65536 - 64, -7 * 65536 + 64))
    pitches += list(range(8 * 65536 - 64, 8 * 65536 + 64))
    for p in pitches:
        assert R.a2_P2I(p) == L.a2o_p2i(p), p


def test_p2i_octaves():
    """One octave up doubles the increment (pitch.c:57-67), and pitch 0 maps to
    2^24 (1.0 in 8:24)."""
    L = ao.lib()
    assert L.a2o_p2i(0) == 1 << 24
    for p in range(-5 * 65536, 5 * 65536, 4099):
        a, b = L.a2o_p2i(p), L.a2o_p2i(p + 65536)
        assert abs(2 * a - b) <= 1


def test_hermite_passes_through_samples():
    """At fractional phase 0 the interpolator returns d[i] (a2_dsp.h:64-74)."""
    L = ao.lib()
    d = (np.sin(np.arange(64) * 0.3) * 30000).astype(np.int16)
    base = d.ctypes.data + 2
    for i in range(1, 60):
        assert L.a2o_hermite(base, i << 8) == int(d[i + 1])


def test_noise_sequence():
    """LCG of a2_dsp.h:37-42 from the default seed (audiality2.h.cmake:62)."""
    L = ao.lib()
    st = C.c_uint32(324357)
    x = 324357
    for _ in range(100):
        x = (x * 1566083941 + 1) & 0xffffffff
        exp = ((x * (x >> 16)) & 0xffffffff) >> 16
        assert L.a2o_noise(C.byref(st)) == exp
        assert st.value == x


def test_wave_preparation_properties():
    """Looped waves are wrapped into the pads, each mip level halves
    (waves.c:59-132)."""
    o = ao.Oracle()
    w = o.builtin_wave("saw")
    prev = None
    for lvl in range(10):
        data, size = o.wave_data(w, lvl)
        assert size == (2048 + (1 << lvl) - 1) >> lvl
        body = data[1:1 + size]
        assert data[0] == body[-1]
        assert np.array_equal(data[1 + size:1 + size + 132],
                              body[np.arange(132) % size])
        prev = body
    o.close()


def test_empty_and_ragged_render():
    """No voices -> silence; buffer sizes that do not divide the frame count
    and are not multiples of 64 render the same samples as buffer 64 where the
    fragment boundaries coincide (core.c:1964-1973)."""
    o = ao.Oracle()
    out = o.render(np.zeros(0, dtype=ao.EVENT_DTYPE), 1000, 96)
    assert out.shape == (1000, 2) and not out.any()
    o.close()
