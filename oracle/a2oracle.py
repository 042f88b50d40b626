"""ctypes binding for oracle/_build/liba2oracle.so (the plain-C port).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(audiality2_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liba2oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

# unit kinds (a2_oracle.h)
WTOSC, PANMIX, FILTER12, WAVESHAPER = 1, 2, 3, 4
FM1, FM2, FM3, FM4, FM3P, FM4P, FM2R, FM4R = 16, 17, 18, 19, 20, 21, 22, 23
EV_WRITE, EV_WAKE, EV_ROOTWRITE, EV_GROUPWRITE = 0, 1, 2, 3


class UnitSpec(C.Structure):
    _fields_ = [("kind", C.c_int), ("ninputs", C.c_int), ("noutputs", C.c_int),
                ("add", C.c_int), ("wireout", C.c_int)]


EVENT_DTYPE = np.dtype([("time", "<u4"), ("kind", "<i4"), ("voice", "<i4"),
                        ("unit", "<i4"), ("reg", "<i4"), ("value", "<i4"),
                        ("dur", "<u4")])


def build(force=False):
    """Compile the port (gcc, a second or two)."""
    if force or not os.path.exists(LIB) or (
            os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "a2_oracle.c"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.a2o_open.restype = C.c_void_p
        L.a2o_open.argtypes = [C.c_int, C.c_int]
        L.a2o_close.argtypes = [C.c_void_p]
        L.a2o_basepitch.argtypes = [C.c_void_p]
        L.a2o_msdur.argtypes = [C.c_void_p]
        L.a2o_msdur.restype = C.c_uint32
        L.a2o_set_noiseseed.argtypes = [C.c_void_p, C.c_uint32]
        L.a2o_builtin_wave.argtypes = [C.c_void_p, C.c_char_p]
        L.a2o_upload_wave.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint,
                                      C.c_void_p, C.c_uint]
        L.a2o_wave_data.restype = C.POINTER(C.c_int16)
        L.a2o_wave_data.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(C.c_uint)]
        L.a2o_new_group.argtypes = [C.c_void_p]
        L.a2o_new_voice.argtypes = [C.c_void_p, C.POINTER(UnitSpec), C.c_int,
                                    C.c_int, C.c_uint, C.c_int]
        L.a2o_kill_voice.argtypes = [C.c_void_p, C.c_int]
        L.a2o_write.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_uint, C.c_uint]
        L.a2o_write_all.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_int, C.c_uint, C.c_uint]
        L.a2o_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_long, C.c_int]
        L.a2o_p2i.restype = C.c_uint
        L.a2o_p2i.argtypes = [C.c_int]
        L.a2o_hermite.argtypes = [C.c_void_p, C.c_uint]
        L.a2o_lerp.argtypes = [C.c_void_p, C.c_uint]
        L.a2o_noise.argtypes = [C.POINTER(C.c_uint32)]
        L.a2o_f12_coeff.argtypes = [C.c_int, C.c_int]
        L.a2o_f12_coeff_array.restype = None
        L.a2o_f12_coeff_array.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def f12_coeff_array(cutoff_values, samplerate):
    """f12_pitch2coeff (filter12.c:65-72) of the port, host libm, for an int32 array."""
    a = np.ascontiguousarray(cutoff_values, dtype=np.int32)
    out = np.empty_like(a)
    lib().a2o_f12_coeff_array(a.ctypes.data, a.size, samplerate, out.ctypes.data)
    return out


class Oracle:
    """Thin OO wrapper over the a2o_* C API."""

    def __init__(self, samplerate=48000, channels=2):
        self.L = lib()
        self.h = self.L.a2o_open(samplerate, channels)
        if not self.h:
            raise MemoryError("a2o_open")
        self.samplerate = samplerate
        self.channels = 1 if channels < 2 else 2

    def close(self):
        if self.h:
            self.L.a2o_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def basepitch(self):
        return self.L.a2o_basepitch(self.h)

    @property
    def msdur(self):
        return self.L.a2o_msdur(self.h)

    def set_noiseseed(self, seed):
        self.L.a2o_set_noiseseed(self.h, seed)

    def builtin_wave(self, name):
        w = self.L.a2o_builtin_wave(self.h, name.encode())
        if w < 0:
            raise ValueError("unknown builtin wave %r" % name)
        return w

    def upload_wave(self, wtype, period, flags, data):
        a = np.ascontiguousarray(data, dtype=np.int16)
        w = self.L.a2o_upload_wave(self.h, wtype, period, flags,
                                   a.ctypes.data, a.size)
        if w < 0:
            raise MemoryError("a2o_upload_wave")
        return w

    def wave_data(self, wave, level):
        n = C.c_uint(0)
        p = self.L.a2o_wave_data(self.h, wave, level, C.byref(n))
        if not p:
            return None
        total = 1 + n.value + 132
        return np.ctypeslib.as_array(p, shape=(total,)).copy(), n.value

    def new_group(self):
        return self.L.a2o_new_group(self.h)

    def group_fbdelay(self, group, regs):
        """{ inline 0 *; fbdelay * *; panmix * > }: the 7 fbdelay registers, 16:16."""
        a = np.ascontiguousarray(regs, dtype=np.int32)
        assert a.size == 7
        self.L.a2o_group_fbdelay.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.L.a2o_group_fbdelay.restype = C.c_int
        if self.L.a2o_group_fbdelay(self.h, group, a.ctypes.data) != 0:
            raise MemoryError("a2o_group_fbdelay")

    def new_voice(self, chain, transpose=0, substart=0, group=-1):
        arr = (UnitSpec * len(chain))(*[UnitSpec(*u) for u in chain])
        v = self.L.a2o_new_voice(self.h, arr, len(chain), transpose, substart,
                                 group)
        if v < 0:
            raise ValueError("a2o_new_voice failed")
        return v

    def write(self, voice, unit, reg, value, start=0, dur=0):
        self.L.a2o_write(self.h, voice, unit, reg, value, start, dur)

    def write_all(self, first, count, unit, reg, values, start=0, dur=0):
        a = np.ascontiguousarray(values, dtype=np.int32)
        self.L.a2o_write_all(self.h, first, count, unit, reg, a.ctypes.data,
                             0 if a.size == 1 else 1, start, dur)

    def render(self, events, frames, buffer=64):
        ev = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        out = np.zeros((frames, self.channels), dtype=np.int32)
        self.L.a2o_render(self.h, ev.ctypes.data if ev.size else None,
                          int(ev.size), out.ctypes.data, frames, buffer)
        return out


def ref_available():
    return os.path.exists(os.path.join(REF_DIR, "a2render"))


def ref_render(a2s_path, program="Song", args=(), samplerate=48000, channels=2,
               buffer=64, frames=4800, noiseseed=None, binary="a2render",
               driver=None, env=None, upload=None, copies=None, cwd=None, shard=None):
    """Run the reference (oracle/_ref/a2render) on a script; returns
    (int32 array [frames, channels], info dict)."""
    import json
    import tempfile
    exe = os.path.join(REF_DIR, binary)
    with tempfile.NamedTemporaryFile(suffix=".raw") as tf:
        cmd = [exe, "-r", str(samplerate), "-b", str(buffer), "-c",
               str(channels), "-n", str(frames), "-p", program, "-o", tf.name]
        for a in args:
            cmd += ["-a", repr(float(a))]
        if noiseseed is not None:
            cmd += ["-s", str(noiseseed)]
        if driver:
            cmd += ["-d", driver]
        if copies:
            cmd += ["-x", str(int(copies))]     # start the program that many times (transposed)
        if shard:
            cmd += ["-X", "%d/%d" % tuple(shard)]   # (rank, world): only copies k with k % world == rank
        if upload:
            # (type, period, flags, length, seed): the harness uploads a pseudo-random wave through
            # a2_UploadWave and passes its handle as the program's last argument
            cmd += ["-U", ":".join(str(int(x)) for x in upload)]
        cmd.append(a2s_path)
        res = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=cwd)
        if res.returncode:
            raise RuntimeError("a2render failed: %s\n%s" % (res.stdout, res.stderr))
        info = json.loads(res.stdout.strip().splitlines()[-1])
        data = np.fromfile(tf.name, dtype="<i4").reshape(-1, channels)
    return data, info
