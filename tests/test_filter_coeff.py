"""filter12's cutoff -> coefficient map (src/units/filter12.c:65-72: float multiply, double `sin`
of the HOST libm) over its whole argument domain.

a2_P2I (src/pitch.c:57-67) sees only the 16 fraction bits of the pitch and (7 - octave) & 31, so
at one sample rate f12_pitch2coeff has 32 x 65536 distinct arguments. The engine tabulates all of
them on the host (same expression, same libm as the reference) and the kernels look the value up:
bit-exact by construction. Here every argument - and a million arbitrary 32-bit ramper values -
goes through the device path and is compared with the port (itself pinned to the reference,
tests/test_oracle.py)."""
import numpy as np
import pytest

from oracle import a2oracle as ao

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rate", [44100, 48000, 22050, 96000])
def test_every_coefficient_argument_matches_host_libm(rate):
    from audiality2_b200 import engine as eng
    shifts = np.arange(32, dtype=np.int64)
    octs = 7 - shifts                                    # a2_P2I shifts by (7 - oct) & 31
    n = np.arange(65536, dtype=np.int64)
    pitch = ((octs[:, None] << 16) | n[None, :]).reshape(-1)          # 2 097 152 pitches, 8:16
    low = np.random.RandomState(1).randint(0, 256, size=pitch.size)   # ramper bits below the pitch
    full = ((pitch << 8) | low).astype(np.int64)
    full = ((full + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)
    rnd = np.random.RandomState(2).randint(-2 ** 31, 2 ** 31, size=1 << 20, dtype=np.int64).astype(np.int32)
    values = np.concatenate([full, rnd])
    e = eng.Engine(rate, 2)
    try:
        dev = e.debug_f12_coeff(values)
    finally:
        e.close()
    ref = ao.f12_coeff_array(values, rate)
    bad = np.nonzero(dev != ref)[0]
    assert bad.size == 0, "%d of %d differ, first: value %d dev %d ref %d" % (
        bad.size, values.size, values[bad[0]], dev[bad[0]], ref[bad[0]])
    assert len(np.unique(ref)) > 100000          # the sweep really exercises the sin() branch
