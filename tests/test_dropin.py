"""GPU tests of the drop-in boundary: the UNMODIFIED reference host
(oracle/_ref/libaudiality2_host.so = reference minus the replaced unit files)
running the golden .a2s scripts on top of our unit plug-in
(audiality2_b200/liba2cu_units.so -> CUDA engine), compared bit for bit with
what the full reference produced (tests/golden/ref_outputs.npz)."""
import os

import numpy as np
import pytest

from cases import CASES, SCRIPTS
from oracle import a2oracle as ao

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_outputs.npz")
HARNESS = os.path.join(ao.REF_DIR, "a2render_cuda")

DROPIN_CASES = sorted(CASES)


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="drop-in harness not built")
@pytest.mark.parametrize("name", DROPIN_CASES)
@pytest.mark.parametrize("driver", ["buffer"])
def test_dropin_matches_reference(name, driver):
    scn = CASES[name]()
    out, info = ao.ref_render(os.path.join(HERE, "golden", name + ".a2s"), "Song",
                              samplerate=scn.samplerate, channels=scn.channels,
                              buffer=scn.buffer, frames=scn.frames,
                              noiseseed=scn.noiseseed,
                              binary="a2render_cuda", driver=driver, upload=scn.uploaded)
    assert info["rt_error"] == 0
    ref = np.load(GOLDEN)[name]
    assert out.shape == ref.shape
    if not np.array_equal(out, ref):
        bad = np.nonzero((out != ref).any(axis=1))[0]
        raise AssertionError("first diff at frame %d (%d differ, max abs %d)" % (
            bad[0], len(bad), np.abs(out.astype(np.int64) - ref).max()))


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="drop-in harness not built")
def test_cuda_driver_registers():
    scn = CASES["osc_pan_ramps"]()
    out, info = ao.ref_render(os.path.join(HERE, "golden", "osc_pan_ramps.a2s"), "Song",
                              frames=scn.frames, binary="a2render_cuda", driver="cuda")
    assert np.array_equal(out, np.load(GOLDEN)["osc_pan_ramps"])


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="drop-in harness not built")
@pytest.mark.parametrize("name", sorted(SCRIPTS))
def test_dropin_song_with_host_units(name):
    """A song whose bus voice chains our inline, the HOST's fbdelay and our
    panmix (download + upload materialisation), leaf voices incl. noise."""
    path, program, frames, rate, buffer = SCRIPTS[name]
    out, info = ao.ref_render(os.path.join(HERE, path), program, samplerate=rate,
                              channels=2, buffer=buffer, frames=frames,
                              binary="a2render_cuda")
    assert info["rt_error"] == 0
    ref = np.load(GOLDEN)["script_" + name]
    if not np.array_equal(out, ref):
        bad = np.nonzero((out != ref).any(axis=1))[0]
        raise AssertionError("first diff at frame %d (%d differ, max abs %d)" % (
            bad[0], len(bad), np.abs(out.astype(np.int64) - ref).max()))
