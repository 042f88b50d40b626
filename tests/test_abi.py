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
    from scenarios import autowire
    for kinds in (["wtosc"], ["wtosc", "panmix"], ["wtosc", "filter12", "panmix"],
                  ["wtosc"] * 8 + ["panmix"], ["fm3", "panmix"], ["fm4r", "panmix"],
                  ["wtosc", "waveshaper", "panmix"]):
        assert engine.Engine.chain_supported(autowire(kinds)), kinds
    assert not engine.Engine.chain_supported([(99, 0, 1, 0, 1)])
