"""
Function-level mirror of the reference's three pybind11 modules (same names, keyword arguments, in/out buffer
conventions and error behaviour), running on libncb200.so:

    simulator.simulate_circuit                     Simulator.cpp:161-212
    simulator_mpi.run_trajectory                   SimulatorMPI.cpp:83-127
    measurement_error_applicator.apply_measurement_error     MeasurementErrorApplicator.cpp:101-122

``simulate_circuit(noisy=True)`` seeds trajectory t with std::mt19937_64(42 + t) exactly like the reference
(Simulator.cpp:124), so for the same inputs it returns the reference's numbers up to floating-point rounding.
Unlike the reference, num_trajectories may exceed 65535 (its loop counter is an unsigned short, Simulator.cpp:123).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .program import PackedCircuit


def _state_buffer(statevector, num_qubits):
    if not isinstance(statevector, np.ndarray) or statevector.dtype != np.complex128 or not statevector.flags.c_contiguous:
        raise TypeError("statevector must be a C-contiguous numpy array of dtype complex128")
    if statevector.size != (1 << num_qubits):
        raise ValueError("statevector must have 2**num_qubits entries")
    return C.c_void_p(statevector.ctypes.data)


def simulate_circuit(instructions, statevector, single_qubit_noise_instructions, two_qubit_noise_instructions,
                     num_qubits, noisy, num_trajectories, return_state, thread_count):
    packed = PackedCircuit(num_qubits, instructions, single_qubit_noise_instructions, two_qubit_noise_instructions, noisy=bool(noisy))
    _lib.check(_lib.lib().ncb_simulate_circuit(C.byref(packed.struct), _state_buffer(statevector, num_qubits), int(bool(noisy)),
                                               int(num_trajectories), int(bool(return_state)), int(thread_count)))


def run_trajectory(instructions, statevector, single_qubit_instructions, two_qubit_instructions, num_qubits, seed, thread_count):
    packed = PackedCircuit(num_qubits, instructions, single_qubit_instructions, two_qubit_instructions, noisy=True)
    _lib.check(_lib.lib().ncb_run_trajectory(C.byref(packed.struct), _state_buffer(statevector, num_qubits), int(seed), int(thread_count)))


def apply_measurement_error(probabilities_array, measurement_noise, qubit_list, num_qubits, num_threads):
    """In place.  Matrices are taken POSITIONALLY from the dictionary (k-th listed qubit <- k-th entry), as in
    MeasurementErrorApplicator.cpp:28-45, 109-112."""
    p = probabilities_array
    if not isinstance(p, np.ndarray) or p.dtype != np.float64 or not p.flags.c_contiguous:
        raise TypeError("probabilities_array must be a C-contiguous numpy array of dtype float64")
    if p.size != (1 << num_qubits):
        raise ValueError("probabilities_array must have 2**num_qubits entries")
    mats = np.ascontiguousarray([np.asarray(m, np.float64) for m in measurement_noise.values()], np.float64)
    qubits = np.asarray([int(q) for q in qubit_list], np.int32)
    if mats.shape[0] < len(qubits):
        raise IndexError("measurement_noise has fewer entries than measured qubits")
    _lib.check(_lib.lib().ncb_apply_measurement_error(C.c_void_p(p.ctypes.data), C.c_void_p(mats.ctypes.data),
                                                      C.c_void_p(qubits.ctypes.data), len(qubits), int(num_qubits), int(num_threads)))


def _matrices_for(measurement_noise: dict, qubit_list):
    """The read-out matrix of every listed qubit: ``measurement_noise[q]`` (the dictionary BuildModel produces is keyed by
    qubit id in ascending order, BuildQubitGateModel.py:843, so for "all qubits in order" this is the reference's positional
    lookup, MeasurementErrorApplicator.cpp:28-45); a dictionary without that key falls back to the k-th entry."""
    items = list(measurement_noise.values())
    out = []
    for k, q in enumerate(qubit_list):
        if int(q) in measurement_noise:
            out.append(np.asarray(measurement_noise[int(q)], np.float64))
        elif k < len(items):
            out.append(np.asarray(items[k], np.float64))
        else:
            raise IndexError(f"no read-out error matrix for qubit {q}")
    return out


def execute_tail_device(probabilities_device: int, num_qubits: int, qubit_list, measurement_noise=None, scale: float = 1.0,
                        stream: int = 0) -> np.ndarray:
    """The tail of QuantumCircuit.execute (QuantumCircuit.py:436-445) on the device: ``probabilities_device`` is a device
    pointer to float64[2^num_qubits] (little-endian, all qubits); returns scale * marginal over ``qubit_list`` ->
    read-out error (``measurement_noise[q]`` on the bit of qubit q, any subset of qubits) -> big-endian order, as a host
    array."""
    qubits = np.asarray([int(q) for q in qubit_list], np.int32)
    out = np.empty(1 << len(qubits), np.float64)
    mats = None
    if measurement_noise:
        mats = np.ascontiguousarray(_matrices_for(measurement_noise, qubits), np.float64)
    _lib.check(_lib.lib().ncb_execute_tail_device(C.c_void_p(int(probabilities_device)), int(num_qubits),
                                                  C.c_void_p(qubits.ctypes.data), len(qubits),
                                                  C.c_void_p(mats.ctypes.data) if mats is not None else None, float(scale),
                                                  C.c_void_p(out.ctypes.data), C.c_void_p(int(stream) or None)))
    return out


class DeviceBuffer:
    """float64 device scratch owned through the C ABI (for callers that do not bring a CUDA allocator)."""

    def __init__(self, count: int):
        p = C.c_void_p()
        _lib.check(_lib.lib().ncb_device_alloc(int(count) * 8, C.byref(p)))
        self.ptr, self.count = int(p.value), int(count)

    def free(self):
        if getattr(self, "ptr", 0):
            _lib.lib().ncb_device_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
