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
PARITY_CASES = [n for n in sorted(CASES) if n != "noise"]


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
