"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a,
loads, and exports every symbol include/a2cu.h declares. No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from audiality2_b200 import build
    path = build.build_engine()
    return C.CDLL(path)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "a2cu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(a2cu_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "liba2cu.so does not export %s" % n


def test_python_binding_covers_header():
    from audiality2_b200 import engine
    assert sorted(engine.SYMBOLS) == declared_symbols()


def test_sass_is_sm100a(lib):
    import subprocess
    from audiality2_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True,
                         text=True).stdout
    assert "sm_100a" in out


def test_open_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from audiality2_b200 import engine
    with pytest.raises(engine.A2cuError):
        engine.Engine()


def test_chain_registry():
    from audiality2_b200 import engine
    from audiality2_b200.chains import autowire
    for kinds in (["wtosc"], ["wtosc", "panmix"], ["wtosc", "filter12", "panmix"],
                  ["wtosc"] * 8 + ["panmix"], ["fm3", "panmix"], ["fm4r", "panmix"],
                  ["wtosc", "waveshaper", "panmix"]):
        assert engine.Engine.chain_supported(autowire(kinds)), kinds
    assert not engine.Engine.chain_supported([(99, 0, 1, 0, 1)])


def test_any_structure_of_replaced_units_has_a_kernel():
    """a2cu_chain_supported (host-side decision, no device needed): structures without a fused kernel
    run on render_generic - two-channel filter12 / waveshaper, adding processors, the bus effects
    inside a leaf chain; only malformed chains and fbdelay (a cooperative bus unit) are refused."""
    from audiality2_b200 import engine
    ok = engine.Engine.chain_supported
    assert ok([(1, 0, 1, 0, 0), (2, 1, 2, 0, 0), (3, 2, 2, 0, 0), (4, 2, 2, 1, 0), (2, 2, 2, 1, 1)])
    assert ok([(18, 0, 1, 0, 0), (1, 0, 1, 1, 0), (3, 1, 1, 0, 0), (2, 1, 1, 1, 1)])
    assert ok([(1, 0, 1, 0, 0), (8, 0, 1, 1, 0), (7, 1, 1, 0, 0), (6, 1, 1, 0, 0), (2, 1, 2, 1, 1)])   # dc, dcblock, limiter
    assert ok([(8, 0, 2, 1, 1)])                                        # struct { dc } on a stereo bus
    assert not ok([(1, 0, 1, 0, 0), (5, 1, 1, 0, 0), (2, 1, 2, 1, 1)])  # fbdelay
    assert not ok([(3, 1, 2, 0, 0), (2, 2, 2, 1, 1)])                   # filter12 must match its I/O
    assert not ok([(1, 1, 1, 0, 0)])                                    # a generator has no inputs
    assert not ok([(1, 0, 1, 0, 0)] * 13)                               # longer than any voice chain


def test_dropin_fails_loudly_without_gpu(tmp_path):
    """The unit plug-in never renders on the CPU: without a CUDA device the host's a2_Open fails
    (our OpenState returns A2_DEVICEOPEN, src/units.c:43-76), it does not fall back."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "oracle", "_ref", "a2render_cuda")
    if not os.path.exists(exe):
        pytest.skip("drop-in harness not built (needs the reference tree)")
    out = tmp_path / "x.raw"
    res = subprocess.run([exe, "-n", "640", "-p", "Song", "-o", str(out),
                          os.path.join(ROOT, "tests", "golden", "osc_pan_ramps.a2s")],
                         capture_output=True, text=True)
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr or "Error opening device" in res.stderr
    assert not out.exists() or out.stat().st_size == 0


def test_units_library_exports_every_descriptor():
    """liba2cu_units.so exports each A2_unitdesc symbol include/a2cu_units.h declares."""
    lib = os.path.join(ROOT, "audiality2_b200", "liba2cu_units.so")
    if not os.path.exists(lib):
        pytest.skip("plug-in not built (needs the reference headers)")
    text = open(os.path.join(ROOT, "include", "a2cu_units.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"extern const struct A2_unitdesc (a2_[a-z0-9]+_unitdesc);", text)
    assert len(names) == 17
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for n in names + ["a2cu_RegisterDriver"]:
        assert (" " + n + "\n") in syms, n


def test_setup_time_clears_and_copies_are_completed():
    """The engine's kernels run on a stream the caller may replace with a NON-BLOCKING one
    (a2cu_set_stream; torch.cuda.Stream is non-blocking). cudaMemset / cudaMemcpy are issued on the
    legacy default stream and may return before the device is done (memset always, memcpy for
    device-to-device and staged pageable copies), and nothing orders them before work on such a
    stream: every creation / growth path must go through memset_done / memcpy_done, which complete
    the operation before returning. Source check (no GPU needed)."""
    import re
    src = open(os.path.join(ROOT, "audiality2_b200", "csrc", "a2cu_engine.cu")).read()
    body = src.replace("cudaError_t r = cudaMemset(p, v, n);", "").replace(
        "cudaError_t r = cudaMemcpy(dst, src, n, kind);", "")
    assert not re.findall(r"(?<![A-Za-z_])cudaMemset\(", body)
    assert not re.findall(r"(?<![A-Za-z_])cudaMemcpy\(", body)
    assert "cudaStreamSynchronize(0)" in src
