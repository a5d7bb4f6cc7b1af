"""
TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (see mcwf_oracle.c for the header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  Nothing under noisycircuits_b200/ does.

Contents
  * ctypes bindings to oracle/_build/liboracle.so (built by `make -C oracle`), a C restatement of the
    reference MCWF path (/root/reference/src/NoisyCircuits/utils/custom/src/*).
  * `density_matrix_probs`: numpy restatement of the density-matrix semantics the reference delegates
    to qiskit-aer (custom/_DensityMatrixSolver.py:43-51, 69-93): per instruction rho <- U rho U^+, then
    rho <- sum_k K_k rho K_k^+ with the gate's Kraus list on the gate's qubits; finally diag(rho).
    qiskit-aer is an unpinned, un-vendored dependency (pyproject.toml:28-43), absent here, so this part
    is "parity unpinned" at the third-party boundary; it is pinned statistically to the compiled
    reference MCWF path by tests/test_gpu_parity.py::test_mcwf_distribution_matches_density_matrix and tests/test_oracle_golden.py (TV <= 3/sqrt(N)).
  * `reference_modules()`: imports the reference's own pybind11 extensions from oracle/_ref when they
    have been built (they are built from /root/reference sources, never copied).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

OPCODES = {"x": 0, "sx": 1, "rz": 2, "rx": 3, "ry": 4, "h": 5, "p": 6,
           "cz": 7, "rzz": 8, "ecr": 9, "cx": 10, "swap": 11, "unitary": 12}
ONE_QUBIT_NOISY = {"x", "sx", "rz", "rx", "ry", "p", "h"}       # FunctionMapper.hpp:55-63
TWO_QUBIT_NOISY = {"cz", "ecr", "rzz", "cx", "swap"}            # FunctionMapper.hpp:64-70


class _Program(C.Structure):
    _fields_ = [("n", C.c_int), ("G", C.c_int),
                ("op", C.c_void_p), ("q0", C.c_void_p), ("q1", C.c_void_p), ("theta", C.c_void_p), ("chan", C.c_void_p),
                ("uni_off", C.c_void_p), ("uni_k", C.c_void_p), ("uni_tq", C.c_void_p), ("uni_pool", C.c_void_p),
                ("ch_off", C.c_void_p), ("ch_cnt", C.c_void_p), ("ch_dim", C.c_void_p), ("kraus_pool", C.c_void_p)]


_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (and oracle/_ref when /root/reference is present) via the Makefile."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_discrete_pick.restype = C.c_int
        _lib.orc_mt64_canonical.restype = C.c_double
        _lib.orc_mt64_next.restype = C.c_uint64
        _lib.orc_noise_step.restype = C.c_int
    return _lib


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Program:
    """Packed form of (instruction list, noise dictionaries) as consumed by mcwf_oracle.c.

    instruction_list: [[name, [q0, q1], param], ...] exactly as HeronDecomposition / EagleDecomposition emit
    it (single-qubit gates carry the qubit twice).  single_qubit_noise / two_qubit_noise: the dictionaries
    QuantumCircuit.__init__ builds (QuantumCircuit.py:158-169).  noisy=False drops every noise step
    (Simulator.cpp:199-203).
    """

    def __init__(self, num_qubits, instruction_list, single_qubit_noise=None, two_qubit_noise=None, noisy=True):
        n, G = int(num_qubits), len(instruction_list)
        self.n, self.G = n, G
        self.op = np.zeros(G, np.int32)
        self.q0 = np.zeros(G, np.int32)
        self.q1 = np.zeros(G, np.int32)
        self.theta = np.zeros(G, np.float64)
        self.chan = np.full(G, -1, np.int32)
        self.uni_off = np.zeros(G, np.int64)
        self.uni_k = np.zeros(G, np.int32)
        self.uni_tq = np.zeros(8 * max(G, 1), np.int32)
        uni_pool, kraus_pool = [], []
        ch_off, ch_cnt, ch_dim = [], [], []
        chan_ids = {}
        uni_len = kraus_len = 0
        sq_items = list(single_qubit_noise.values()) if single_qubit_noise else []
        for g, (name, qubits, param) in enumerate(instruction_list):
            name = str(name)
            self.op[g] = OPCODES[name]
            qubits = [int(q) for q in qubits]
            if name == "unitary":
                U = np.ascontiguousarray(np.asarray(param, dtype=np.complex128))
                tq = qubits[::-1]                                   # Simulator.cpp:182
                self.uni_k[g] = len(tq)
                self.uni_tq[8 * g: 8 * g + len(tq)] = tq
                self.uni_off[g] = uni_len
                uni_pool.append(U.reshape(-1)); uni_len += U.size
                self.q0[g] = tq[0]; self.q1[g] = tq[-1]
                continue
            self.q0[g] = qubits[0]
            self.q1[g] = qubits[1] if len(qubits) > 1 else qubits[0]
            self.theta[g] = float(param) if param is not None else 0.0
            if not noisy:
                continue
            if name in ONE_QUBIT_NOISY:
                key = ("1q", qubits[0], name)
                if key not in chan_ids:
                    ops = sq_items[qubits[0]][name]               # positional lookup, NoiseParser.hpp:52-71, 132
                    chan_ids[key] = len(ch_off)
                    ch_off.append(kraus_len); ch_cnt.append(len(ops)); ch_dim.append(2)
                    for K in ops:
                        kraus_pool.append(np.asarray(K, np.complex128).reshape(-1)); kraus_len += 4
                self.chan[g] = chan_ids[key]
            elif name in TWO_QUBIT_NOISY:
                key = ("2q", (qubits[0], qubits[1]), name)
                if key not in chan_ids:
                    ops = two_qubit_noise[name][(qubits[0], qubits[1])]   # NoiseParser.hpp:135-137 (.at -> IndexError)
                    chan_ids[key] = len(ch_off)
                    ch_off.append(kraus_len); ch_cnt.append(len(ops)); ch_dim.append(4)
                    for K in ops:
                        kraus_pool.append(np.asarray(K, np.complex128).reshape(-1)); kraus_len += 16
                self.chan[g] = chan_ids[key]
        self.uni_pool = np.concatenate(uni_pool) if uni_pool else np.zeros(1, np.complex128)
        self.kraus_pool = np.concatenate(kraus_pool) if kraus_pool else np.zeros(1, np.complex128)
        self.ch_off = np.asarray(ch_off if ch_off else [0], np.int64)
        self.ch_cnt = np.asarray(ch_cnt if ch_cnt else [0], np.int32)
        self.ch_dim = np.asarray(ch_dim if ch_dim else [2], np.int32)
        self.kraus_count = np.array([self.ch_cnt[c] if c >= 0 else 0 for c in self.chan], np.int32)
        self._c = _Program(n, G, _ptr(self.op), _ptr(self.q0), _ptr(self.q1), _ptr(self.theta), _ptr(self.chan),
                           _ptr(self.uni_off), _ptr(self.uni_k), _ptr(self.uni_tq), _ptr(self.uni_pool),
                           _ptr(self.ch_off), _ptr(self.ch_cnt), _ptr(self.ch_dim), _ptr(self.kraus_pool))

    # -- entry points ------------------------------------------------------------------------------------------
    def reference_uniforms(self, seed: int) -> np.ndarray:
        """u[g] the reference would draw for `std::mt19937_64(seed)` (0.5 where no draw is consumed)."""
        out = np.empty(self.G, np.float64)
        lib().orc_reference_uniforms(C.c_uint64(seed), C.c_int(self.G), _ptr(self.kraus_count), _ptr(out))
        return out

    def trajectory(self, uniforms, mask_fix=True, diagnostics=False):
        """Final normalised state of one trajectory driven by per-instruction uniforms."""
        u = np.ascontiguousarray(uniforms, np.float64)
        assert u.shape == (self.G,)
        state = np.empty(1 << self.n, np.complex128)
        choices = np.empty(self.G, np.int32)
        pc = np.empty(self.G, np.float64)
        lib().orc_trajectory(C.byref(self._c), _ptr(u), C.c_int(int(mask_fix)), _ptr(state), _ptr(choices), _ptr(pc))
        return (state, choices, pc) if diagnostics else state

    def mcwf(self, num_trajectories, uniforms=None, seed0=42, mask_fix=True, threads=1) -> np.ndarray:
        """Mean |psi|^2 over trajectories; uniforms=None reproduces the reference's seed0+t MT streams."""
        probs = np.empty(1 << self.n, np.float64)
        up = C.c_void_p(None)
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, np.float64)
            assert uniforms.shape == (num_trajectories, self.G)
            up = _ptr(uniforms)
        lib().orc_mcwf(C.byref(self._c), C.c_int64(num_trajectories), up, C.c_uint64(seed0), C.c_int(int(mask_fix)),
                       C.c_int(threads), _ptr(probs))
        return probs

    def pure_state(self, mask_fix=True) -> np.ndarray:
        state = np.empty(1 << self.n, np.complex128)
        lib().orc_pure_state(C.byref(self._c), C.c_int(int(mask_fix)), _ptr(state))
        return state


def measurement_error(probs, measurement_noise: dict, qubit_list, num_qubits):
    """In-place restatement of measurement_error_applicator.apply_measurement_error (positional matrices)."""
    mats = np.ascontiguousarray([np.asarray(m, np.float64) for m in measurement_noise.values()], np.float64)
    q = np.asarray(list(qubit_list), np.int32)
    assert probs.dtype == np.float64 and probs.flags.c_contiguous
    lib().orc_measurement_error(_ptr(probs), C.c_int(num_qubits), _ptr(mats), _ptr(q), C.c_int(len(q)))
    return probs


# ---------------------------------------------------------------------------------------------------------------
# Density-matrix restatement (numpy)
# ---------------------------------------------------------------------------------------------------------------
def gate_matrix(name: str, theta: float = 0.0) -> np.ndarray:
    """Matrices of the 12 named gates; 2q matrices are indexed bit(qubits[0]) + 2*bit(qubits[1])
    (QuantumGates.hpp:209-219), i.e. exactly what the C++ kernels compute."""
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    r2 = 1 / np.sqrt(2)
    if name == "x":   return np.array([[0, 1], [1, 0]], complex)
    if name == "sx":  return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
    if name == "rz":  return np.diag([c - 1j * s, c + 1j * s])
    if name == "rx":  return np.array([[c, -1j * s], [-1j * s, c]])
    if name == "ry":  return np.array([[c, -s], [s, c]], complex)
    if name == "h":   return r2 * np.array([[1, 1], [1, -1]], complex)
    if name == "p":   return np.diag([1, np.cos(theta) + 1j * np.sin(theta)])
    if name == "cz":  return np.diag([1, 1, 1, -1]).astype(complex)
    if name == "rzz": return np.diag([c - 1j * s, c + 1j * s, c + 1j * s, c - 1j * s])
    if name == "ecr": return r2 * np.array([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]])
    if name == "cx":  return np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], complex)
    if name == "swap":return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], complex)
    raise KeyError(name)


def _apply_left(rho, M, qubits, n):
    """rho <- M rho on `qubits` (matrix bit b <-> qubits[b]); rho has 2n axes, row qubit q = axis n-1-q."""
    k = len(qubits)
    Mt = M.reshape([2] * (2 * k))                       # [out_{k-1}..out_0, in_{k-1}..in_0]
    axes = [n - 1 - q for q in reversed(qubits)]          # in_{k-1} <-> qubits[k-1] ...
    out = np.tensordot(Mt, rho, axes=(list(range(k, 2 * k)), axes))
    return np.moveaxis(out, list(range(k)), axes)


def _apply_right_dagger(rho, M, qubits, n):
    """rho <- rho M^+ on the column index of `qubits`."""
    k = len(qubits)
    Mt = M.conj().reshape([2] * (2 * k))
    axes = [2 * n - 1 - q for q in reversed(qubits)]
    out = np.tensordot(Mt, rho, axes=(list(range(k, 2 * k)), axes))
    return np.moveaxis(out, list(range(k)), axes)


def density_matrix(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise) -> np.ndarray:
    n = int(num_qubits)
    rho = np.zeros([2] * (2 * n), complex)
    rho[(0,) * (2 * n)] = 1.0
    sq_items = list(single_qubit_noise.values()) if single_qubit_noise else []
    for name, qubits, param in instruction_list:
        qubits = [int(q) for q in qubits]
        if name == "unitary":
            U = np.asarray(param, complex)
            tq = qubits[::-1]
            rho = _apply_right_dagger(_apply_left(rho, U, tq, n), U, tq, n)
            continue
        if name in ONE_QUBIT_NOISY:
            tq, ops = [qubits[0]], (sq_items[qubits[0]][name] if sq_items else None)
        else:
            tq, ops = [qubits[0], qubits[1]], (two_qubit_noise[name][(qubits[0], qubits[1])] if two_qubit_noise else None)
        U = gate_matrix(name, 0.0 if param is None else float(param))
        rho = _apply_right_dagger(_apply_left(rho, U, tq, n), U, tq, n)
        if ops is not None:
            acc = np.zeros_like(rho)
            for K in ops:
                K = np.asarray(K, complex)
                if not K.any():
                    continue
                acc += _apply_right_dagger(_apply_left(rho, K, tq, n), K, tq, n)
            rho = acc
    return rho.reshape(1 << n, 1 << n)


def marginal_little_endian(probs, num_qubits, keep):
    """Marginal over the kept qubits, little-endian in the kept qubits' ascending order
    (what diag(partial_trace(rho, trace_qubits)) gives, custom/_DensityMatrixSolver.py:90-93)."""
    n = int(num_qubits)
    keep = sorted(int(q) for q in keep)
    if len(keep) == n:
        return np.ascontiguousarray(probs, np.float64)
    t = np.asarray(probs).reshape([2] * n)
    axes = tuple(n - 1 - q for q in range(n) if q not in keep)
    return np.ascontiguousarray(t.sum(axis=axes).reshape(-1), np.float64)


def density_matrix_probs(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise, qubits=None):
    rho = density_matrix(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise)
    probs = np.ascontiguousarray(np.diag(rho).real, np.float64)
    if qubits is None:
        return probs
    return marginal_little_endian(probs, num_qubits, qubits)


# ---------------------------------------------------------------------------------------------------------------
# The compiled reference (oracle/_ref), when present
# ---------------------------------------------------------------------------------------------------------------
def reference_modules():
    """(simulator, simulator_mpi, measurement_error_applicator) built from the reference's own sources,
    or None when oracle/_ref has not been built."""
    ref = os.path.join(_HERE, "_ref")
    if not os.path.isdir(ref) or not any(f.startswith("simulator.") for f in os.listdir(ref)):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import importlib
    try:
        return tuple(importlib.import_module(m) for m in ("simulator", "simulator_mpi", "measurement_error_applicator"))
    except ImportError:
        return None
