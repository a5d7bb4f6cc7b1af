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
