"""Python host binding of the voice-render engine (ctypes over include/a2cu.h).

This mirrors the C ABI one to one; nothing is computed in Python.  The product
path is CUDA only: importing works anywhere, but `Engine()` raises when
liba2cu.so is missing or no CUDA device is usable - there is no CPU fallback
(and this package never imports anything from oracle/).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# A2CU_LIB: an alternative build of the same library (kernel A/B experiments, profiles/scripts/build_variant.sh)
LIB_PATH = os.environ.get("A2CU_LIB") or os.path.join(HERE, "liba2cu.so")

# unit kinds / wave types (include/a2cu.h)
WTOSC, PANMIX, FILTER12, WAVESHAPER = 1, 2, 3, 4
FM1, FM2, FM3, FM4, FM3P, FM4P, FM2R, FM4R = 16, 17, 18, 19, 20, 21, 22, 23
WOFF, WNOISE, WWAVE, WMIPWAVE = 0, 1, 2, 3
LOOPED = 0x100


class A2cuError(RuntimeError):
    pass


class UnitSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ninputs", C.c_int32),
                ("noutputs", C.c_int32), ("add", C.c_int32),
                ("wireout", C.c_int32)]


_lib = None

# name -> (restype, argtypes); the complete export list of include/a2cu.h
_VP, _I, _U, _U64, _I32P = C.c_void_p, C.c_int, C.c_uint, C.c_uint64, C.POINTER(C.c_int32)
SYMBOLS = {
    "a2cu_open": (_VP, [_I, _I, _I]),
    "a2cu_close": (None, [_VP]),
    "a2cu_last_error": (C.c_char_p, []),
    "a2cu_set_stream": (_I, [_VP, _VP]),
    "a2cu_basepitch": (_I, [_VP]),
    "a2cu_msdur": (C.c_uint32, [_VP]),
    "a2cu_now": (_U64, [_VP]),
    "a2cu_set_root_wake_period": (_I, [_VP, C.c_uint32]),
    "a2cu_set_noiseseed": (_I, [_VP, C.c_uint32]),
    "a2cu_set_noise_state_ptr": (_I, [_VP, C.POINTER(C.c_uint32)]),
    "a2cu_wave_builtin": (_I, [_VP, C.c_char_p]),
    "a2cu_wave_upload": (_I, [_VP, _I, _U, _U, _VP, _U]),
    "a2cu_wave_upload_prepared": (_I, [_VP, _I, _U, _U, _VP, _VP]),
    "a2cu_wave_unload": (_I, [_VP, _I]),
    "a2cu_wave_read": (_I, [_VP, _I, _I, _VP, _U, C.POINTER(C.c_uint)]),
    "a2cu_group_new": (_I, [_VP]),
    "a2cu_chain_supported": (_I, [C.POINTER(UnitSpec), _I]),
    "a2cu_bank_new": (_I, [_VP, C.POINTER(UnitSpec), _I, _I, _VP, _VP, _U]),
    "a2cu_bank_kill": (_I, [_VP, _I, _I, _U64]),
    "a2cu_bank_write": (_I, [_VP, _I, _I, _I, _I, C.c_int32, _U64, C.c_uint32]),
    "a2cu_bank_write_all": (_I, [_VP, _I, _I, _I, _VP, _I, _U64, C.c_uint32]),
    "a2cu_bank_wake": (_I, [_VP, _I, _I, _U64]),
    "a2cu_bank_enable": (_I, [_VP, _I, _I]),
    "a2cu_group_write": (_I, [_VP, _I, _I, C.c_int32, _U64, C.c_uint32]),
    "a2cu_root_write": (_I, [_VP, _I, C.c_int32, _U64, C.c_uint32]),
    "a2cu_run": (_I, [_VP, _U, _U, _VP]),
    "a2cu_run_async": (_I, [_VP, _U, _U, _VP]),
    "a2cu_master_devptr": (_VP, [_VP]),
    "a2cu_sync": (_I, [_VP]),
    "a2cu_submit": (_I, [_VP, _U, _U]),
    "a2cu_submit_dev": (_I, [_VP, _U, _U, _VP]),
    "a2cu_collect": (_I, [_VP, _I, _VP]),
    "a2cu_set_post_root_stage": (_I, [_VP, _I]),
    "a2cu_set_output_format": (_I, [_VP, _I]),
    "a2cu_apply_root_stage": (_I, [_VP, _VP, _VP, _U, _U, _U64]),
    "a2cu_xchg_create": (_I, [_VP, _I, _I, _U, _U, _VP]),
    "a2cu_xchg_connect_ipc": (_I, [_VP, _VP]),
    "a2cu_xchg_connect_local": (_I, [_VP, C.POINTER(_VP)]),
    "a2cu_xchg_enable": (_I, [_VP, _I]),
    "a2cu_xchg_close": (_I, [_VP]),
    "a2cu_launch_count": (_U64, [_VP]),
    "a2cu_split_launch_count": (_U64, [_VP]),
    "a2cu_set_split": (_I, [_VP, _I]),
    "a2cu_split_profile": (_I, [_VP, _I, C.POINTER(C.c_uint64)]),
    "a2cu_split_trace": (_I, [_VP, C.POINTER(C.c_uint64)]),
    "a2cu_split_trace_reset": (_I, [_VP]),
    "a2cu_bank_kernel_name": (C.c_char_p, [_VP, _I]),
    "a2cu_bank_state_bytes": (_I, [_VP, _I]),
    "a2cu_last_render_ms": (C.c_float, [_VP]),
    "a2cu_last_mix_ms": (C.c_float, [_VP]),
    "a2cu_h2d_bytes": (_U64, [_VP]),
    "a2cu_d2h_bytes": (_U64, [_VP]),
    "a2cu_set_timing": (_I, [_VP, _I]),
    "a2cu_debug_f12_coeff": (_I, [_VP, _VP, _I, _VP]),
    # drop-in ("block") mode, used by the unit plug-in
    "a2cu_pool_open": (_I, [_VP, C.POINTER(UnitSpec), _I]),
    "a2cu_pool_alloc": (_I, [_VP, _I]),
    "a2cu_pool_free": (_I, [_VP, _I, _I]),
    "a2cu_block_begin": (_I, [_VP]),
    "a2cu_block_bus": (_I, [_VP]),
    "a2cu_block_init": (_I, [_VP, _I, _I, _I, _I, _U, _U]),
    "a2cu_block_write": (_I, [_VP, _I, _I, _I, _I, C.c_int32, _I, _U, _U, C.c_uint32]),
    "a2cu_block_proc": (_I, [_VP, _I, _I, _U, _U, _I]),
    "a2cu_pm_alloc": (_I, [_VP]),
    "a2cu_pm_free": (_I, [_VP, _I]),
    "a2cu_block_pm_write": (_I, [_VP, _I, _I, C.c_int32, _U, C.c_uint32]),
    "a2cu_block_pm_proc": (_I, [_VP, _I, _I, _I, _I, _I, _I, _U, _U]),
    "a2cu_block_flush": (_I, [_VP]),
    "a2cu_block_upload": (_I, [_VP, _I, _I, _U, _U, _VP]),
    "a2cu_block_download": (_I, [_VP, _I, _I, _U, _U, _VP, _I]),
    # generic units (one replaced unit per call) and per-voice command runs
    "a2cu_block_run": (_U64, [_VP, _I, _U64]),
    "a2cu_unit_alloc": (_I, [_VP, _I, _I, _I]),
    "a2cu_unit_free": (_I, [_VP, _I]),
    "a2cu_block_unit_init": (_I, [_VP, _I, _I, _U]),
    "a2cu_block_unit_write": (_I, [_VP, _I, _I, C.c_int32, _I, _U, C.c_uint32]),
    "a2cu_block_unit_proc": (_I, [_VP, _I, _I, _I, _I, _I, _U, _U]),
    "a2cu_block_bus_add": (_I, [_VP, _I, _I, _U, _U]),
}


def load_library(path=LIB_PATH):
    """dlopen liba2cu.so and type every exported entry point. Raises if the
    CUDA extension has not been built - the product never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(path):
            raise A2cuError("%s not built: run `python -m audiality2_b200.build` "
                            "(CUDA extension is mandatory, no CPU fallback)" % path)
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _specs(chain):
    return (UnitSpec * len(chain))(*[UnitSpec(*u) for u in chain])


class Engine:
    """a2cu_engine handle. Times are 24:8 frames since open, values 16:16."""

    def __init__(self, samplerate=48000, channels=2, device=0):
        self.L = load_library()
        self.h = self.L.a2cu_open(device, samplerate, channels)
        if not self.h:
            raise A2cuError("a2cu_open failed: %s" % self.L.a2cu_last_error().decode())
        self.samplerate = samplerate
        self.channels = 1 if channels < 2 else 2
        self.post_root = True
        self.out_dtype = np.int32

    def close(self):
        if getattr(self, "h", None):
            self.L.a2cu_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r < 0:
            raise A2cuError("a2cu error %d: %s" % (r, self.L.a2cu_last_error().decode()))
        return r

    # -- info
    @property
    def basepitch(self):
        return self.L.a2cu_basepitch(self.h)

    @property
    def msdur(self):
        return self.L.a2cu_msdur(self.h)

    @property
    def now(self):
        return self.L.a2cu_now(self.h)

    @property
    def launches(self):
        return self.L.a2cu_launch_count(self.h)

    @property
    def split_launches(self):
        return self.L.a2cu_split_launch_count(self.h)

    def set_split(self, on):
        self._ck(self.L.a2cu_set_split(self.h, int(on)))

    def split_profile(self, enable=True, read=True):
        out = (C.c_uint64 * 8)()
        self._ck(self.L.a2cu_split_profile(self.h, int(enable), out if read else None))
        return list(out)

    def split_trace(self):
        """[role][fragment][begin/end] cycles of CTA 0's pipeline in the last render_split launch."""
        out = (C.c_uint64 * 768)()
        self._ck(self.L.a2cu_split_trace(self.h, out))
        return np.array(list(out), dtype=np.int64).reshape(6, 64, 2)

    def split_trace_reset(self):
        self._ck(self.L.a2cu_split_trace_reset(self.h))

    def set_stream(self, cuda_stream):
        self._ck(self.L.a2cu_set_stream(self.h, cuda_stream))

    def set_root_wake_period(self, period):
        self._ck(self.L.a2cu_set_root_wake_period(self.h, period))

    def set_noiseseed(self, seed):
        self._ck(self.L.a2cu_set_noiseseed(self.h, seed))

    def set_timing(self, on=True):
        self._ck(self.L.a2cu_set_timing(self.h, int(on)))

    def last_render_ms(self):
        return float(self.L.a2cu_last_render_ms(self.h))

    def last_mix_ms(self):
        return float(self.L.a2cu_last_mix_ms(self.h))

    @property
    def h2d_bytes(self):
        return self.L.a2cu_h2d_bytes(self.h)

    @property
    def d2h_bytes(self):
        return self.L.a2cu_d2h_bytes(self.h)

    def debug_f12_coeff(self, cutoff_values):
        a = np.ascontiguousarray(cutoff_values, dtype=np.int32)
        out = np.empty_like(a)
        self._ck(self.L.a2cu_debug_f12_coeff(self.h, a.ctypes.data, a.size, out.ctypes.data))
        return out

    def set_output_format(self, fmt):
        """'i32' (8:24, default), 'f32' (v / 2^23, the audio drivers' edge) or 'i16' (v >> 8, the
        wave writer's): converted inside the root stage."""
        code = {"i32": 0, "f32": 1, "i16": 2}[fmt]
        self._ck(self.L.a2cu_set_output_format(self.h, code))
        self.out_dtype = {0: np.int32, 1: np.float32, 2: np.int16}[code]

    def set_post_root_stage(self, on):
        self.post_root = bool(on)
        self._ck(self.L.a2cu_set_post_root_stage(self.h, int(on)))

    # -- waves
    def builtin_wave(self, name):
        return self._ck(self.L.a2cu_wave_builtin(self.h, name.encode()))

    def upload_wave(self, wtype, period, flags, data):
        a = np.ascontiguousarray(data, dtype=np.int16)
        return self._ck(self.L.a2cu_wave_upload(self.h, wtype, period, flags,
                                                a.ctypes.data, a.size))

    def unload_wave(self, wave):
        self._ck(self.L.a2cu_wave_unload(self.h, wave))

    def wave_data(self, wave, level):
        n = C.c_uint(0)
        cap = (1 << 24) + 256
        cnt = self._ck(self.L.a2cu_wave_read(self.h, wave, level, None, 0, C.byref(n)))
        buf = np.zeros(1 + n.value + 132, dtype=np.int16)
        cnt = self._ck(self.L.a2cu_wave_read(self.h, wave, level, buf.ctypes.data,
                                             buf.size, C.byref(n)))
        return buf[:cnt], n.value

    # -- structure
    def new_group(self):
        return self._ck(self.L.a2cu_group_new(self.h))

    @staticmethod
    def chain_supported(chain):
        return bool(load_library().a2cu_chain_supported(_specs(chain), len(chain)))

    def new_bank(self, chain, nvoices, transpose=None, group=None, substart=0):
        tr = None if transpose is None else np.ascontiguousarray(transpose, dtype=np.int32)
        gr = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        return self._ck(self.L.a2cu_bank_new(
            self.h, _specs(chain), len(chain), nvoices,
            None if tr is None else tr.ctypes.data,
            None if gr is None else gr.ctypes.data, substart))

    def bank_kernel_name(self, bank):
        return self.L.a2cu_bank_kernel_name(self.h, bank).decode()

    def bank_state_bytes(self, bank):
        return self._ck(self.L.a2cu_bank_state_bytes(self.h, bank))

    # -- control
    def write(self, bank, voice, unit, reg, value, when=None, dur=0):
        when = self.now if when is None else when
        self._ck(self.L.a2cu_bank_write(self.h, bank, voice, unit, reg, value, when, dur))

    def write_all(self, bank, unit, reg, values, when=None, dur=0):
        when = self.now if when is None else when
        a = np.ascontiguousarray(values, dtype=np.int32)
        stride = 0 if a.size == 1 else 1
        self._ck(self.L.a2cu_bank_write_all(self.h, bank, unit, reg, a.ctypes.data,
                                            stride, when, dur))

    def bank_enable(self, bank, on=True):
        self._ck(self.L.a2cu_bank_enable(self.h, bank, int(on)))

    def wake(self, bank, voice, when):
        self._ck(self.L.a2cu_bank_wake(self.h, bank, voice, when))

    def kill(self, bank, voice, when):
        self._ck(self.L.a2cu_bank_kill(self.h, bank, voice, when))

    def group_write(self, group, reg, value, when=None, dur=0):
        when = self.now if when is None else when
        self._ck(self.L.a2cu_group_write(self.h, group, reg, value, when, dur))

    def root_write(self, reg, value, when=None, dur=0):
        when = self.now if when is None else when
        self._ck(self.L.a2cu_root_write(self.h, reg, value, when, dur))

    # -- render
    def run(self, frames, buffer=64):
        """a2_Run() analogue: returns int32 8:24 [frames, channels] (host)."""
        ch = self.channels if self.post_root else 2
        out = np.empty((frames, ch), dtype=self.out_dtype if self.post_root else np.int32)
        self._ck(self.L.a2cu_run(self.h, frames, buffer, out.ctypes.data))
        return out

    def submit(self, frames, buffer=64):
        """Queue one window without waiting (a2cu_submit); returns a ticket."""
        self._frames_of = getattr(self, "_frames_of", {})
        t = self._ck(self.L.a2cu_submit(self.h, frames, buffer))
        self._frames_of[t] = frames
        return t

    def submit_dev(self, frames, buffer, dev_ptr):
        """Queue one window whose output stays in device memory (a2cu_submit_dev)."""
        return self._ck(self.L.a2cu_submit_dev(self.h, frames, buffer, dev_ptr))

    def collect_spans(self, ticket):
        """Wait for a submit_dev window's kernels; latches last_render_ms / last_mix_ms."""
        self._ck(self.L.a2cu_collect(self.h, ticket, None))

    def collect(self, ticket, out=None):
        """Wait for a submitted window and return its int32 8:24 output."""
        if out is None:
            ch = self.channels if self.post_root else 2
            out = np.empty((self._frames_of[ticket], ch), dtype=self.out_dtype if self.post_root else np.int32)
        self._ck(self.L.a2cu_collect(self.h, ticket, out.ctypes.data))
        return out

    def run_async(self, frames, buffer=64, dev_ptr=None):
        self._ck(self.L.a2cu_run_async(self.h, frames, buffer, dev_ptr))

    def master_devptr(self):
        return self.L.a2cu_master_devptr(self.h)

    def sync(self):
        self._ck(self.L.a2cu_sync(self.h))

    # -- multi-GPU: root-bus exchange inside the render kernel (a2cu_xchg_*)
    def xchg_create(self, rank, world, max_frames, timeout_ms=0):
        """Allocate this rank's symmetric buffer; returns its 64-byte IPC handle."""
        h = (C.c_ubyte * 64)()
        self._ck(self.L.a2cu_xchg_create(self.h, rank, world, max_frames, timeout_ms, h))
        return bytes(h)

    def xchg_connect_ipc(self, handles):
        """handles: world x 64 bytes in rank order (other processes' buffers)."""
        buf = b"".join(handles)
        self._ck(self.L.a2cu_xchg_connect_ipc(self.h, buf))

    def xchg_connect_local(self, engines):
        """engines[r] = the Engine of rank r, all living in this process."""
        arr = (_VP * len(engines))(*[e.h for e in engines])
        self._ck(self.L.a2cu_xchg_connect_local(self.h, arr))

    def xchg_enable(self, on=True):
        self._ck(self.L.a2cu_xchg_enable(self.h, int(on)))

    def xchg_close(self):
        self._ck(self.L.a2cu_xchg_close(self.h))

    def apply_root_stage(self, dev_rootbus, dev_master, frames, buffer=64):
        self._ck(self.L.a2cu_apply_root_stage(self.h, dev_rootbus, dev_master,
                                              frames, buffer, 0))
