"""In-tree build of the CUDA libraries (sm_100a only).

    python -m audiality2_b200.build            # liba2cu.so (engine + kernels)

nvcc cross-compiles without a GPU; the resulting .so files are git-ignored but
travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "liba2cu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-cudart", "shared", "-Xfatbin", "-compress-all",
]
# translation units of liba2cu.so: host + bus kernels, and one per kernel family
UNITS = ["a2cu_engine.cu", "a2cu_reg_bank_wt.cu", "a2cu_reg_bank_fm.cu", "a2cu_reg_split.cu"]
OBJDIR = os.path.join(HERE, "build")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_engine(force=False, verbose=False):
    """nvcc -c every unit (in parallel, only the stale ones), then link."""
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "a2cu.h"))
    os.makedirs(OBJDIR, exist_ok=True)
    procs, objs = [], []
    for u in UNITS:
        src = os.path.join(CSRC, u)
        obj = os.path.join(OBJDIR, u[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, hdrs + [src]):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
                "-I", os.path.join(ROOT, "include"), "-c", "-o", obj, src]
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait():
            raise subprocess.CalledProcessError(p.returncode, cmd)
    if procs or not os.path.exists(LIB):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared",
                               "-o", LIB] + objs)
    return LIB


def build_plugin():
    """Drop-in unit library + its check harness. They compile against the
    reference's headers where they lie, so this only works where
    /root/reference exists (the dev container); the .so files travel."""
    if not os.path.isdir("/root/reference/src"):
        return None
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "plugin")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
    return os.path.join(HERE, "liba2cu_units.so")


def build_all(force=False, verbose=False):
    return [build_engine(force, verbose), build_plugin()]


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
