"""ctypes binding of libncb200.so (C ABI: include/ncb200.h).  There is no CPU fallback: if the library is missing
the import of the solver classes fails loudly, and every run function fails when no CUDA device is usable."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libncb200.so")

NCB_OK = 0
MODE_MCWF, MODE_DENSITY, MODE_PURE = 0, 1, 2
RNG_PHILOX, RNG_MT19937_REFERENCE, RNG_EXPLICIT = 0, 1, 2
(Q_NUM_QUBITS, Q_STATE_BITS, Q_TILE_BITS, Q_REG_BITS, Q_NUM_SEGMENTS, Q_NUM_PASSES, Q_NUM_OPS, Q_NUM_DRAWS,
 Q_NORM_PRESERVING, Q_NUM_INSTRUCTIONS, Q_RESIDENT) = range(11)


class NcbCircuit(C.Structure):
    _fields_ = [("num_qubits", C.c_int32), ("num_instructions", C.c_int32),
                ("opcode", C.c_void_p), ("q0", C.c_void_p), ("q1", C.c_void_p), ("theta", C.c_void_p),
                ("channel", C.c_void_p), ("unitary_offset", C.c_void_p), ("unitary_nq", C.c_void_p),
                ("unitary_qubits", C.c_void_p), ("unitary_pool", C.c_void_p),
                ("num_channels", C.c_int32), ("channel_offset", C.c_void_p), ("channel_count", C.c_void_p),
                ("channel_dim", C.c_void_p), ("kraus_pool", C.c_void_p)]


class NcbOptions(C.Structure):
    _fields_ = [("tile_bits", C.c_int32), ("reg_bits", C.c_int32), ("low_bits", C.c_int32), ("fuse", C.c_int32),
                ("device", C.c_int32), ("strict_order", C.c_int32), ("independent_trajectories", C.c_int32),
                ("specialize", C.c_int32)]


class NcbRunParams(C.Structure):
    _fields_ = [("traj_begin", C.c_int64), ("traj_count", C.c_int64), ("rng", C.c_int32), ("probs_on_device", C.c_int32),
                ("seed", C.c_uint64), ("uniforms", C.c_void_p), ("probs", C.c_void_p), ("accumulate", C.c_int32),
                ("batch", C.c_int32), ("states", C.c_void_p), ("stream", C.c_void_p),
                ("out_sweeps", C.c_int64), ("out_events", C.c_int64), ("out_hard_events", C.c_int64),
                ("out_launches", C.c_int64), ("out_kernel_ms", C.c_double), ("out_tile_launches", C.c_int64),
                ("out_h2d_bytes", C.c_int64), ("out_d2h_bytes", C.c_int64), ("out_spec_launches", C.c_int64)]


# every symbol include/ncb200.h declares
EXPORTS = ["ncb_last_error", "ncb_version", "ncb_program_create", "ncb_program_destroy", "ncb_program_update", "ncb_program_describe",
           "ncb_program_query", "ncb_program_order", "ncb_program_events", "ncb_program_reference_uniforms", "ncb_mcwf_run", "ncb_mcwf_launch",
           "ncb_mcwf_collect", "ncb_pure_run",
           "ncb_density_run", "ncb_apply_measurement_error", "ncb_apply_measurement_error_device",
           "ncb_execute_tail_device", "ncb_device_alloc", "ncb_device_free", "ncb_program_spec_build", "ncb_program_spec_source",
           "ncb_simulate_circuit", "ncb_run_trajectory"]

_lib = None


class NcbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libncb200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
                "or make -C noisycircuits_b200/csrc).  noisycircuits_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.ncb_last_error.restype = C.c_char_p
        L.ncb_version.restype = C.c_int
        L.ncb_program_create.argtypes = [C.POINTER(NcbCircuit), C.c_int, C.POINTER(NcbOptions), C.POINTER(C.c_void_p)]
        L.ncb_program_destroy.argtypes = [C.c_void_p]
        L.ncb_program_destroy.restype = None
        L.ncb_program_update.argtypes = [C.c_void_p, C.POINTER(NcbCircuit)]
        L.ncb_program_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.ncb_program_describe.restype = C.c_int64
        L.ncb_program_query.argtypes = [C.c_void_p, C.c_int]
        L.ncb_program_query.restype = C.c_int64
        L.ncb_program_order.argtypes = [C.c_void_p, C.c_void_p]
        L.ncb_program_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.ncb_program_events.restype = C.c_int64
        L.ncb_program_reference_uniforms.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.ncb_mcwf_run.argtypes = [C.c_void_p, C.POINTER(NcbRunParams)]
        L.ncb_mcwf_launch.argtypes = [C.c_void_p, C.POINTER(NcbRunParams)]
        L.ncb_mcwf_collect.argtypes = [C.c_void_p, C.POINTER(NcbRunParams)]
        L.ncb_pure_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ncb_density_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ncb_apply_measurement_error.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32]
        L.ncb_apply_measurement_error_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.ncb_execute_tail_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.ncb_device_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p)]
        L.ncb_device_free.argtypes = [C.c_void_p]
        L.ncb_program_spec_build.argtypes = [C.c_void_p]
        L.ncb_program_spec_source.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int64]
        L.ncb_program_spec_source.restype = C.c_int64
        L.ncb_simulate_circuit.argtypes = [C.POINTER(NcbCircuit), C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_uint32]
        L.ncb_run_trajectory.argtypes = [C.POINTER(NcbCircuit), C.c_void_p, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != NCB_OK:
        raise NcbError(rc, lib().ncb_last_error().decode("utf-8", "replace"))
