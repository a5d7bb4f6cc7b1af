"""
Solver plug-in for NoisyCircuits: the three classes a ``sim_backend`` module must export
(/root/reference/src/NoisyCircuits/utils/custom/__init__.py:5-7), with the ``custom`` back-end's signatures, running on
the B200 engine (libncb200.so).

    RemoteExecutor(num_qubits, single_qubit_noise, two_qubit_noise, num_cores).run(num_trajectories, instruction_list)
        -> float64[2^n], little-endian, all qubits, before measurement error, divided by N
           (custom/_ParallelExecutor.py:18-65; call site QuantumCircuit.py:411-420)
    DensityMatrixSolver(num_qubits, single_qubit_noise, two_qubit_noise, instruction_list, num_cores).solve(qubits)
        -> float64[2^len(qubits)], little-endian over the kept qubits (custom/_DensityMatrixSolver.py:29-93)
    PureStateSolver(num_qubits, instruction_list, num_cores, return_statevector).solve()
        -> float64[2^n] probabilities or complex128[2^n] state (custom/_PureStateSolver.py:28-66)

``num_cores`` is accepted for signature compatibility and ignored (the GPU replaces the OpenMP threads).
Extra keyword arguments (all optional) expose what the CPU path cannot do: ``rng`` ("philox" counter-based draws,
the default; "reference" reproduces the reference's mt19937_64(42 + t) streams so results equal the reference's to
rounding), ``seed``, and trajectory sharding for multi-GPU runs.
"""
from __future__ import annotations

import numpy as np

from . import _lib, cache
from .program import PackedCircuit, Program

_RNG = {"philox": _lib.RNG_PHILOX, "reference": _lib.RNG_MT19937_REFERENCE}


def _marginal_little_endian(probs: np.ndarray, num_qubits: int, keep) -> np.ndarray:
    """Marginal over ``keep`` (little-endian in ascending kept-qubit order), i.e. diag(partial_trace(rho, rest))
    (custom/_DensityMatrixSolver.py:90-93); same arithmetic as HelperFunctions.py:9-32."""
    keep = sorted(int(q) for q in keep)
    if len(keep) == num_qubits:
        return np.require(probs, dtype=np.float64, requirements=["C"])
    t = probs.reshape([2] * num_qubits)
    axes = tuple(num_qubits - 1 - q for q in range(num_qubits) if q not in keep)
    return np.require(t.sum(axis=axes).reshape(-1), dtype=np.float64, requirements=["C"])


def _fingerprint(instruction_list):
    """Content key of an instruction list (compared as a whole: no hash collisions)."""
    items = []
    for name, qubits, param in instruction_list:
        if isinstance(param, np.ndarray):
            param = param.tobytes()
        items.append((name, tuple(qubits), param))
    return tuple(items)


class RemoteExecutor:
    """MCWF trajectories on the GPU (mirror of custom/_ParallelExecutor.py:14-65)."""

    def __init__(self, num_qubits: int, single_qubit_noise: dict, two_qubit_noise: dict, num_cores: int = 2,
                 rng: str = "philox", seed: int = 42, **engine_options) -> None:
        if rng not in _RNG:
            raise ValueError(f"rng must be one of {sorted(_RNG)}")
        self.num_qubits = num_qubits
        self.single_qubit_noise = single_qubit_noise
        self.two_qubit_noise = two_qubit_noise
        self.num_cores = num_cores
        self.rng = rng
        self.seed = seed
        if rng == "reference":
            # the reference consumes its mt19937_64 stream in program order and every branch decision sees the state of
            # that order (Simulator.cpp:76-94): no re-ordering of commuting instructions in this mode
            engine_options.setdefault("strict_order", True)
        self.engine_options = engine_options
        self.last_stats = None
        self._cache, self._cache_digest = (None, None), None
        self._skeleton = (None, None)

    # Generated kernels (ncb_options.specialize) cost a few seconds of compilation per new kernel set.  Without an explicit
    # specialize=... in the engine options: from SPECIALIZE_FROM trajectories per call the "parametric" kernels are used
    # (coefficients as a kernel parameter: one kernel set per gate SKELETON, so a parameter sweep compiles once), and calls
    # so long that the compilation is noise -- trajectories x 2^n >= LITERAL_FROM_AMPLITUDES, e.g. 65536 trajectories of 20
    # qubits -- get the literal-coefficient kernels (10-35 % faster, one kernel set per parameter set; the on-disk cache
    # keeps the 64 most recently used sets).
    SPECIALIZE_FROM = 4096
    LITERAL_FROM_AMPLITUDES = 1 << 36

    def _auto_specialize(self, traj_count: int):
        if traj_count < self.SPECIALIZE_FROM:
            return False
        return True if (int(traj_count) << self.num_qubits) >= self.LITERAL_FROM_AMPLITUDES else "parametric"

    def _program(self, instruction_list, traj_count: int = 0) -> Program:
        opts = dict(self.engine_options)
        opts.setdefault("specialize", self._auto_specialize(traj_count))
        # fast path: the same list object contents and the same noise dictionaries as last time (in-place edits of the
        # dictionaries are not seen here -- call invalidate()); otherwise the packed bytes decide (cache.get_program)
        key = (_fingerprint(instruction_list), str(opts["specialize"]), id(self.single_qubit_noise), id(self.two_qubit_noise))
        # (the cache may have recycled the object for another circuit meanwhile: its digest says what it holds now)
        if self._cache[0] != key or getattr(self._cache[1], "digest", None) != self._cache_digest:
            prog = cache.get_program(self._pack(key[0]) or self._pack(key[0], instruction_list), _lib.MODE_MCWF, **opts)
            self._cache, self._cache_digest = (key, prog), prog.digest
        return self._cache[1]

    def _pack(self, fingerprint, instruction_list=None):
        """Packed circuit of an instruction list.  A training loop's next circuit is the last one's gate skeleton with other
        parameters: then only the parameter array is replaced (``PackedCircuit.with_thetas``; no Python loop over the
        instructions, no re-packing of the Kraus operators).  Without ``instruction_list``: None when the skeleton is new."""
        skeleton = (tuple((name, qubits) for name, qubits, _p in fingerprint), id(self.single_qubit_noise), id(self.two_qubit_noise))
        if instruction_list is None:
            known, base = self._skeleton
            if known != skeleton or any(name == "unitary" for name, _q, _p in fingerprint):
                return None
            thetas = np.zeros(base.theta.shape[0])
            thetas[:len(fingerprint)] = [0.0 if p is None else float(p) for _n, _q, p in fingerprint]
            return base.with_thetas(thetas)
        packed = PackedCircuit(self.num_qubits, instruction_list, self.single_qubit_noise, self.two_qubit_noise, noisy=True)
        self._skeleton = (skeleton, packed)
        return packed

    def invalidate(self) -> None:
        """Forget the compiled program (after editing the noise dictionaries in place)."""
        self._cache, self._cache_digest = (None, None), None
        self._skeleton = (None, None)

    def run_sum(self, instruction_list, traj_begin: int, traj_count: int, uniforms=None, device_probs=None,
                accumulate=False, stream=0):
        """Unnormalised sum of |psi_t|^2 over trajectories [traj_begin, traj_begin + traj_count): the building block
        of sharded (multi-GPU) runs; trajectory ids key the RNG so the result does not depend on the sharding."""
        prog = self._program(instruction_list, traj_count)
        seed = self.seed
        probs, stats = prog.run_mcwf(traj_count, traj_begin=traj_begin, rng=_RNG[self.rng], seed=seed, uniforms=uniforms,
                                     device_probs=device_probs, accumulate=accumulate, stream=stream)
        self.last_stats = stats
        return probs

    def run(self, num_trajectories: int, instruction_list: list) -> np.ndarray:
        probs = self.run_sum(instruction_list, 0, num_trajectories)
        return probs / float(num_trajectories)

    def run_many(self, num_trajectories: int, instruction_lists) -> list:
        """``run`` for a sequence of circuits (a QNN parameter sweep over one gate skeleton).  A few engine
        objects take turns: while the GPU runs circuit i in one, host threads pack the next circuits and compile them INTO the
        others in place (``Program.update``: device buffers, streams and -- for one gate skeleton -- the loaded kernels
        are reused).  Same results as calling ``run`` in a loop (trajectory ids 0..N-1 for every circuit)."""
        lists = list(instruction_lists)
        if not lists:
            return []
        makers = ((lambda instr=instr: PackedCircuit(self.num_qubits, instr, self.single_qubit_noise, self.two_qubit_noise, noisy=True))
                  for instr in lists)
        return self._run_packed(num_trajectories, makers, len(lists))

    def run_sweep(self, num_trajectories: int, instruction_list: list, thetas) -> list:
        """``run`` for the gate skeleton ``instruction_list`` with every row of ``thetas`` ([sets][instructions]: the
        parameter of each instruction; entries of parameter-free gates are ignored) -- the parameter sets of a training
        loop as a batch: the skeleton is packed once, each set only replaces the angle array."""
        thetas = np.asarray(thetas, np.float64)
        if thetas.ndim != 2 or thetas.shape[1] != len(instruction_list):
            raise ValueError("thetas must have shape [sets][len(instruction_list)]")
        base = PackedCircuit(self.num_qubits, instruction_list, self.single_qubit_noise, self.two_qubit_noise, noisy=True)
        pad = np.zeros(base.theta.shape[0] - thetas.shape[1])
        makers = ((lambda row=row: base.with_thetas(np.concatenate([row, pad]))) for row in thetas)
        return self._run_packed(num_trajectories, makers, thetas.shape[0])

    PREPARE_THREADS = 2          # host threads that pack / compile the next circuits of a sweep while the GPU runs one

    def _run_packed(self, num_trajectories, makers, count):
        import threading
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        opts = dict(self.engine_options)
        opts.setdefault("specialize", "parametric" if num_trajectories >= self.SPECIALIZE_FROM else False)   # a sweep: one kernel set
        makers = iter(makers)
        depth = max(1, int(self.PREPARE_THREADS))
        # the engine objects live as long as the executor: a training loop calls this every step, and creating an engine
        # means a dozen device allocations (each can stall for tens of milliseconds next to a large state batch)
        key = (depth,) + tuple(sorted((k, str(v)) for k, v in opts.items()))
        # engines: one in flight on the GPU, one launched behind it, `depth` being prepared
        n_eng = depth + 2
        if getattr(self, "_sweep_key", None) != key:
            self._sweep_key, self._sweep_engines = key, [None] * n_eng
        engines = self._sweep_engines
        lock = threading.Lock()
        taken = [0]

        def prepare():
            # circuit idx goes into engine idx % n_eng; the caller submits a preparation only when that engine's previous
            # circuit (idx - n_eng) has been collected
            with lock:
                idx = taken[0]
                taken[0] += 1
                make = next(makers)
            packed = make()
            slot = idx % n_eng
            if engines[slot] is None:
                engines[slot] = Program(packed, _lib.MODE_MCWF, **opts)
            else:
                try:
                    engines[slot].update(packed)
                except _lib.NcbError:                      # another geometry: a new engine object
                    engines[slot] = Program(packed, _lib.MODE_MCWF, **opts)
            return idx, engines[slot]

        # circuit i is launched (ncb_mcwf_launch: programs whose planning happens on the device return at once) before circuit
        # i - 1 is collected, so the GPU always has the next circuit's kernels queued behind -- or, the trunk, beside -- the
        # current one's; a preparation is only submitted once the engine it will overwrite has been collected
        out, ready, futs, inflight = [], {}, deque(), None
        with ThreadPoolExecutor(max_workers=depth) as pool:
            for _ in range(min(depth + 1, count)):
                futs.append(pool.submit(prepare))
            submitted = len(futs)

            def collect(item):
                prog, handle = item
                probs, stats = prog.collect_mcwf(handle)
                self.last_stats = stats
                out.append(probs / float(num_trajectories))

            try:
                for i in range(count):
                    while i not in ready:
                        idx, prog = futs.popleft().result()
                        ready[idx] = prog
                    prog = ready.pop(i)
                    handle = prog.launch_mcwf(num_trajectories, traj_begin=0, rng=_RNG[self.rng], seed=self.seed)
                    previous, inflight = inflight, (prog, handle)
                    if previous is not None:
                        collect(previous)
                    if submitted < count:                    # (goes into the engine of circuit i - 1, collected just now)
                        futs.append(pool.submit(prepare))
                        submitted += 1
                last, inflight = inflight, None
                if last is not None:
                    collect(last)
            except BaseException:
                # leave no run in flight and no half-prepared engine behind: the next sweep starts with fresh engine objects
                if inflight is not None:
                    try:
                        inflight[0].collect_mcwf(inflight[1])
                    except Exception:
                        pass
                for f in futs:
                    f.cancel()
                self._sweep_key, self._sweep_engines = None, []
                raise
        return out


class DensityMatrixSolver:
    """Density-matrix evolution on the vectorised 4^n state (mirror of custom/_DensityMatrixSolver.py:29-93; the
    reference delegates the arithmetic to qiskit-aer, semantics in SURVEY.md 3.3)."""

    def __init__(self, num_qubits: int, single_qubit_noise: dict, two_qubit_noise: dict, instruction_list: list,
                 num_cores: int = 1, **engine_options) -> None:
        self.num_qubits = num_qubits
        self.instruction_list = instruction_list
        self.num_cores = num_cores
        packed = PackedCircuit(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise, noisy=True)
        engine_options.setdefault("reg_bits", 4)
        self._program = Program(packed, _lib.MODE_DENSITY, **engine_options)

    def solve(self, qubits: list) -> np.ndarray:
        probs = self._program.run_density()
        return _marginal_little_endian(probs, self.num_qubits, qubits)


class PureStateSolver:
    """Noise-free state-vector run (mirror of custom/_PureStateSolver.py:28-66)."""

    def __init__(self, num_qubits: int, instruction_list: list, num_cores: int = 1, return_statevector: bool = False,
                 **engine_options) -> None:
        self.num_qubits = num_qubits
        self.instruction_list = instruction_list
        self.num_cores = num_cores
        self.return_statevector = return_statevector
        packed = PackedCircuit(num_qubits, instruction_list, noisy=False)
        engine_options.setdefault("reg_bits", 4)
        self._program = Program(packed, _lib.MODE_PURE, **engine_options)

    def solve(self) -> np.ndarray:
        if self.return_statevector:
            state, _ = self._program.run_pure(want_state=True, want_probs=False)
            return state
        _, probs = self._program.run_pure(want_state=False, want_probs=True)
        return probs
