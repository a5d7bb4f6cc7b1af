"""
Packing of the reference's Python-side inputs (instruction list + noise dictionaries) into the C ABI's
``ncb_circuit`` and a thin object wrapper around a compiled ``ncb_program``.

What is packed is exactly what the reference's native layer unpacks on every call
(/root/reference/src/NoisyCircuits/utils/custom/src/Simulator.cpp:162-196, NoiseParser.hpp:52-140,
FunctionMapper.hpp:25-79): gate name -> opcode, single-qubit gates carry their qubit twice, ``unitary`` gates
carry a matrix and get no noise, single-qubit Kraus lists are looked up by POSITION of the qubit's entry in the
dictionary (NoiseParser.hpp:52-71, 132), two-qubit lists by ``(gate, (q0, q1))`` (NoiseParser.hpp:135-137; a missing
entry is an IndexError there, as here).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

OPCODES = {"x": 0, "sx": 1, "rz": 2, "rx": 3, "ry": 4, "h": 5, "p": 6,
           "cz": 7, "rzz": 8, "ecr": 9, "cx": 10, "swap": 11, "unitary": 12}
ONE_QUBIT_NOISY = frozenset({"x", "sx", "rz", "rx", "ry", "p", "h"})      # FunctionMapper.hpp:55-63
TWO_QUBIT_NOISY = frozenset({"cz", "ecr", "rzz", "cx", "swap"})           # FunctionMapper.hpp:64-70


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class PackedCircuit:
    """numpy-backed ``ncb_circuit``.  Keeps the arrays alive for as long as the struct is in use."""

    def __init__(self, num_qubits, instruction_list, single_qubit_noise=None, two_qubit_noise=None, noisy=True):
        if not isinstance(num_qubits, (int, np.integer)) or num_qubits < 1:
            raise ValueError("num_qubits must be a positive integer")
        n, G = int(num_qubits), len(instruction_list)
        self.num_qubits, self.num_instructions = n, G
        self.opcode = np.zeros(max(G, 1), np.int32)
        self.q0 = np.zeros(max(G, 1), np.int32)
        self.q1 = np.zeros(max(G, 1), np.int32)
        self.theta = np.zeros(max(G, 1), np.float64)
        self.channel = np.full(max(G, 1), -1, np.int32)
        self.unitary_offset = np.zeros(max(G, 1), np.int64)
        self.unitary_nq = np.zeros(max(G, 1), np.int32)
        self.unitary_qubits = np.zeros(8 * max(G, 1), np.int32)
        uni_pool, kraus_pool = [], []
        ch_off, ch_cnt, ch_dim = [], [], []
        chan_ids = {}
        uni_len = kraus_len = 0
        sq_items = list(single_qubit_noise.values()) if single_qubit_noise else []
        for g, entry in enumerate(instruction_list):
            name, qubits, param = entry[0], entry[1], entry[2]
            name = str(name)
            if name not in OPCODES:
                raise ValueError(f"instruction {g}: unknown gate '{name}'")
            self.opcode[g] = OPCODES[name]
            qubits = [int(q) for q in qubits]
            if name == "unitary":
                U = np.ascontiguousarray(np.asarray(param, dtype=np.complex128))
                k = len(qubits)
                if U.shape != (1 << k, 1 << k):
                    raise ValueError(f"instruction {g}: unitary of shape {U.shape} on {k} qubits")
                if k > 8:
                    raise ValueError(f"instruction {g}: unitary on more than 8 qubits")
                tq = qubits[::-1]                                         # Simulator.cpp:182
                self.unitary_nq[g] = k
                self.unitary_qubits[8 * g: 8 * g + k] = tq
                self.unitary_offset[g] = uni_len
                uni_pool.append(U.reshape(-1)); uni_len += U.size
                self.q0[g] = tq[0]; self.q1[g] = tq[-1]
                continue
            self.q0[g] = qubits[0]
            self.q1[g] = qubits[1] if len(qubits) > 1 else qubits[0]
            self.theta[g] = 0.0 if param is None else float(param)
            if not noisy:
                continue
            if name in ONE_QUBIT_NOISY:
                key = ("1q", qubits[0], name)
                if key not in chan_ids:
                    try:
                        ops = sq_items[qubits[0]][name]
                    except (IndexError, KeyError) as e:
                        raise IndexError(f"no single-qubit noise entry for gate '{name}' on qubit {qubits[0]}") from e
                    chan_ids[key] = len(ch_off)
                    ch_off.append(kraus_len); ch_cnt.append(len(ops)); ch_dim.append(2)
                    for K in ops:
                        K = np.asarray(K, np.complex128)
                        if K.shape != (2, 2):
                            raise ValueError(f"single-qubit Kraus operator of shape {K.shape}")
                        kraus_pool.append(K.reshape(-1)); kraus_len += 4
                self.channel[g] = chan_ids[key]
            elif name in TWO_QUBIT_NOISY:
                key = ("2q", (qubits[0], qubits[1]), name)
                if key not in chan_ids:
                    try:
                        ops = two_qubit_noise[name][(qubits[0], qubits[1])]
                    except (KeyError, TypeError) as e:
                        raise IndexError(f"no two-qubit noise entry for gate '{name}' on qubits {(qubits[0], qubits[1])}") from e
                    chan_ids[key] = len(ch_off)
                    ch_off.append(kraus_len); ch_cnt.append(len(ops)); ch_dim.append(4)
                    for K in ops:
                        K = np.asarray(K, np.complex128)
                        if K.shape != (4, 4):
                            raise ValueError(f"two-qubit Kraus operator of shape {K.shape}")
                        kraus_pool.append(K.reshape(-1)); kraus_len += 16
                self.channel[g] = chan_ids[key]
        self.unitary_pool = np.concatenate(uni_pool) if uni_pool else np.zeros(1, np.complex128)
        self.kraus_pool = np.concatenate(kraus_pool) if kraus_pool else np.zeros(1, np.complex128)
        self.num_channels = len(ch_off)
        self.channel_offset = np.asarray(ch_off if ch_off else [0], np.int64)
        self.channel_count = np.asarray(ch_cnt if ch_cnt else [0], np.int32)
        self.channel_dim = np.asarray(ch_dim if ch_dim else [2], np.int32)
        self.struct = _lib.NcbCircuit(
            n, G, _ptr(self.opcode), _ptr(self.q0), _ptr(self.q1), _ptr(self.theta), _ptr(self.channel),
            _ptr(self.unitary_offset), _ptr(self.unitary_nq), _ptr(self.unitary_qubits), _ptr(self.unitary_pool),
            self.num_channels, _ptr(self.channel_offset), _ptr(self.channel_count), _ptr(self.channel_dim),
            _ptr(self.kraus_pool))

    def with_thetas(self, thetas) -> "PackedCircuit":
        """The same circuit (gate skeleton, qubits, Kraus lists) with other gate parameters: thetas[g] replaces the
        parameter of instruction g (entries of parameter-free gates are ignored by the compiler).  No Python loop over the
        instructions: what a parameter sweep needs per circuit."""
        import copy
        other = copy.copy(self)
        other.theta = np.ascontiguousarray(thetas, np.float64).copy()
        if other.theta.shape != self.theta.shape:
            raise ValueError(f"thetas must have shape {self.theta.shape}")
        other.struct = _lib.NcbCircuit(
            self.num_qubits, self.num_instructions, _ptr(other.opcode), _ptr(other.q0), _ptr(other.q1), _ptr(other.theta), _ptr(other.channel),
            _ptr(other.unitary_offset), _ptr(other.unitary_nq), _ptr(other.unitary_qubits), _ptr(other.unitary_pool),
            other.num_channels, _ptr(other.channel_offset), _ptr(other.channel_count), _ptr(other.channel_dim), _ptr(other.kraus_pool))
        return other

    @property
    def kraus_count(self) -> np.ndarray:
        """Operators in each instruction's Kraus list (0: no noise step)."""
        return np.array([self.channel_count[c] if c >= 0 else 0 for c in self.channel[:self.num_instructions]], np.int32)


class Program:
    """A compiled program (``ncb_program``).  Compilation is host-only; the first run uploads it to the GPU."""

    def __init__(self, packed: PackedCircuit, mode: int = _lib.MODE_MCWF, tile_bits: int = 0, reg_bits: int = 0,
                 low_bits: int = -1, fuse: int = 1, device: int = -1, strict_order: bool = False,
                 independent_trajectories: bool = False, specialize: bool = False):
        self._h = C.c_void_p(None)
        self.packed = packed
        self.mode = mode
        opts = _lib.NcbOptions(tile_bits, reg_bits, low_bits, fuse, device, int(bool(strict_order)),
                               int(bool(independent_trajectories)), 2 if specialize in (2, "parametric") else int(bool(specialize)))
        _lib.check(_lib.lib().ncb_program_create(C.byref(packed.struct), mode, C.byref(opts), C.byref(self._h)))
        self.num_qubits = packed.num_qubits
        self.num_instructions = packed.num_instructions

    def update(self, packed: "PackedCircuit") -> None:
        """Replaces the circuit in place (``ncb_program_update``): same mode and options, the engine's device buffers and --
        for ``specialize="parametric"`` programs of the same gate skeleton -- its loaded kernels are kept."""
        _lib.check(_lib.lib().ncb_program_update(self._h, C.byref(packed.struct)))
        self.packed = packed
        self.num_qubits = packed.num_qubits
        self.num_instructions = packed.num_instructions

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().ncb_program_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, what: int) -> int:
        return int(_lib.lib().ncb_program_query(self._h, what))

    def describe(self) -> str:
        n = _lib.lib().ncb_program_describe(self._h, None, 0)
        buf = C.create_string_buffer(int(n))
        _lib.lib().ncb_program_describe(self._h, buf, n)
        return buf.value.decode()

    def order(self) -> np.ndarray:
        """order[pos] = index in the instruction list of the instruction executed at position pos."""
        out = np.zeros(max(self.num_instructions, 1), np.int32)
        _lib.check(_lib.lib().ncb_program_order(self._h, _ptr(out)))
        return out[:self.num_instructions]

    def events(self, uniforms) -> np.ndarray:
        """Jump events (instr, kmin, kmax) of the trajectory driven by ``uniforms`` (host-only)."""
        u = np.ascontiguousarray(uniforms, np.float64)
        assert u.shape == (self.num_instructions,)
        cap = max(self.num_instructions, 1)
        out = np.zeros((cap, 3), np.int32)
        n = _lib.lib().ncb_program_events(self._h, _ptr(u), _ptr(out), cap)
        return out[:n]

    def specialize_build(self) -> bool:
        """Generates and compiles this program's specialised kernels into the cubin cache (host only, a few seconds; the
        first run of a program created with ``specialize=True`` does it implicitly).  False: nothing to specialise or
        nvcc not usable -- the interpreting kernels are used."""
        return _lib.lib().ncb_program_spec_build(self._h) == 0

    def specialized_source(self, seg_begin: int = 0, seg_end: int = 1 << 30) -> str:
        n = _lib.lib().ncb_program_spec_source(self._h, int(seg_begin), int(min(seg_end, 1 << 30)), None, 0)
        buf = C.create_string_buffer(int(n))
        _lib.lib().ncb_program_spec_source(self._h, int(seg_begin), int(min(seg_end, 1 << 30)), buf, n)
        return buf.value.decode()

    def reference_uniforms(self, seed: int) -> np.ndarray:
        u = np.empty(max(self.num_instructions, 1), np.float64)
        _lib.check(_lib.lib().ncb_program_reference_uniforms(self._h, C.c_uint64(seed), _ptr(u)))
        return u[:self.num_instructions]

    def run_mcwf(self, traj_count, traj_begin=0, rng=_lib.RNG_PHILOX, seed=42, uniforms=None, return_states=False,
                 batch=0, stream=0, device_probs=None, accumulate=False, _launch_only=False):
        """Sum over trajectories [traj_begin, traj_begin + traj_count) of |psi|^2 (NOT divided by the count).

        Returns (probs, stats) or (probs, states, stats).  ``device_probs``: a device pointer (int) to float64[2^n]
        that receives (or, with accumulate, is incremented by) the sum instead of a host array.
        """
        dim = 1 << self.num_qubits
        rp = _lib.NcbRunParams()
        rp.traj_begin, rp.traj_count, rp.rng, rp.seed = int(traj_begin), int(traj_count), int(rng), int(seed)
        rp.batch, rp.stream = int(batch), C.c_void_p(int(stream) or None)
        rp.accumulate = int(bool(accumulate))
        keep = []
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, np.float64)
            if u.shape != (traj_count, self.num_instructions):
                raise ValueError(f"uniforms must have shape {(traj_count, self.num_instructions)}")
            keep.append(u)
            rp.uniforms = _ptr(u)
            rp.rng = _lib.RNG_EXPLICIT
        probs = None
        if device_probs is not None:
            rp.probs, rp.probs_on_device = C.c_void_p(int(device_probs)), 1
        else:
            probs = np.zeros(dim, np.float64)
            rp.probs = _ptr(probs)
        states = None
        if return_states:
            states = np.empty((traj_count, dim), np.complex128)
            rp.states = _ptr(states)
        if _launch_only:
            _lib.check(_lib.lib().ncb_mcwf_launch(self._h, C.byref(rp)))
            return rp, probs, keep                      # (everything the pending run points into)
        _lib.check(_lib.lib().ncb_mcwf_run(self._h, C.byref(rp)))
        stats = self._stats(rp)
        return (probs, states, stats) if return_states else (probs, stats)

    @staticmethod
    def _stats(rp):
        return {"sweeps": rp.out_sweeps, "events": rp.out_events, "hard_events": rp.out_hard_events,
                "launches": rp.out_launches, "kernel_ms": rp.out_kernel_ms, "tile_launches": rp.out_tile_launches,
                "h2d_bytes": rp.out_h2d_bytes, "d2h_bytes": rp.out_d2h_bytes, "specialized_launches": rp.out_spec_launches}

    def launch_mcwf(self, traj_count, traj_begin=0, rng=_lib.RNG_PHILOX, seed=42, stream=0, device_probs=None, accumulate=False):
        """First half of ``run_mcwf`` (``ncb_mcwf_launch``): enqueues the run -- on a stream of this program's own when
        ``stream`` is 0 -- and returns a handle for ``collect_mcwf``.  Programs whose state fits one CTA (generated kernels,
        Philox draws) return without waiting for the GPU, so a caller can keep several programs in flight; the others have
        finished their kernels on return."""
        return self.run_mcwf(traj_count, traj_begin=traj_begin, rng=rng, seed=seed, stream=stream, device_probs=device_probs,
                             accumulate=accumulate, _launch_only=True)

    def collect_mcwf(self, handle):
        """Second half (``ncb_mcwf_collect``): waits for the run ``launch_mcwf`` started and returns (probs, stats)."""
        rp, probs, _keep = handle
        _lib.check(_lib.lib().ncb_mcwf_collect(self._h, C.byref(rp)))
        return probs, self._stats(rp)

    def run_pure(self, want_state=True, want_probs=False, stream=0):
        dim = 1 << self.num_qubits
        state = np.empty(dim, np.complex128) if want_state else None
        probs = np.empty(dim, np.float64) if want_probs else None
        _lib.check(_lib.lib().ncb_pure_run(self._h, _ptr(state) if want_state else None,
                                           _ptr(probs) if want_probs else None, C.c_void_p(int(stream) or None)))
        return state, probs

    def run_density(self, stream=0) -> np.ndarray:
        probs = np.empty(1 << self.num_qubits, np.float64)
        _lib.check(_lib.lib().ncb_density_run(self._h, _ptr(probs), C.c_void_p(int(stream) or None)))
        return probs
