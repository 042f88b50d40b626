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
