"""TEST INFRASTRUCTURE ONLY: ctypes face of tests/emu/emu.cpp (CPU emulation of the tile kernel; see its header)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libncb_emu.so")
SRCS = [os.path.join(HERE, "emu.cpp"), os.path.join(ROOT, "noisycircuits_b200", "csrc", "ncb_compile.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "noisycircuits_b200", "csrc", f) for f in ("ncb_common.h", "ncb_compile.h", "ncb_schedule.h")]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO] + SRCS, check=True)
        L = C.CDLL(SO)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_order.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_query.argtypes = [C.c_void_p, C.c_int]
        L.emu_query.restype = C.c_longlong
        L.emu_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
        L.emu_describe.restype = C.c_longlong
        L.emu_mcwf_states.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_run_deterministic.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_last_error.restype = C.c_char_p
        L.emu_philox_uniform.restype = C.c_double
        L.emu_philox_uniform.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint]
        _lib = L
    return _lib


class EmuProgram:
    def __init__(self, packed, mode=0, tile_bits=12, reg_bits=4, low_bits=3, fuse=1, reorder=1):
        self.packed = packed
        self.h = lib().emu_create(C.byref(packed.struct), mode, tile_bits, reg_bits, low_bits, fuse, reorder)
        if not self.h:
            raise RuntimeError(lib().emu_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().emu_destroy(self.h)
            self.h = None

    def order(self):
        out = np.zeros(max(self.packed.num_instructions, 1), np.int32)
        lib().emu_order(self.h, out.ctypes.data)
        return out[:self.packed.num_instructions]

    def query(self, what):
        return int(lib().emu_query(self.h, what))

    def describe(self):
        n = lib().emu_describe(self.h, None, 0)
        buf = C.create_string_buffer(int(n))
        lib().emu_describe(self.h, buf, n)
        return buf.value.decode()

    def mcwf_states(self, uniforms, share=False):
        u = np.ascontiguousarray(uniforms, np.float64)
        count = u.shape[0]
        nbits = self.query(6)
        out = np.empty((count, 1 << nbits), np.complex128)
        stats = np.zeros(4, np.int64)
        if lib().emu_mcwf_states(self.h, count, u.ctypes.data, out.ctypes.data, stats.ctypes.data, int(share)) != 0:
            raise RuntimeError(lib().emu_last_error().decode())
        return out, {"events": int(stats[0]), "hard_events": int(stats[1]), "sweeps": int(stats[2]), "steps": int(stats[3])}

    def run_deterministic(self):
        nbits = self.query(6)
        out = np.empty(1 << nbits, np.complex128)
        if lib().emu_run_deterministic(self.h, out.ctypes.data) != 0:
            raise RuntimeError(lib().emu_last_error().decode())
        return out
