"""
The ``execute`` call chains of the reference around the solver plug-in, and the registration of this package as a
NoisyCircuits ``sim_backend``.

    execute          QuantumCircuit.execute        (/root/reference/src/NoisyCircuits/QuantumCircuit.py:353-446)
    execute_sharded  QuantumCircuitMPI.execute     (/root/reference/src/NoisyCircuits/QuantumCircuitMPI.py:263-368)
                     with torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) in place of mpi4py:
                     contiguous trajectory shards (:329-332), one sum-reduction of float64[2^n] (:352), then the same
                     post-processing on the reduced vector (:355-366).
"""
from __future__ import annotations

import sys

import numpy as np

from . import simulator
from .solvers import RemoteExecutor, _marginal_little_endian


def shard_range(num_trajectories: int, rank: int, world_size: int):
    """Contiguous split of QuantumCircuitMPI.py:329-332: (local_start, local_count)."""
    base, rem = divmod(int(num_trajectories), int(world_size))
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def postprocess(probs: np.ndarray, num_qubits: int, qubits, measurement_error: dict, num_cores: int = 1,
                apply_measurement_error=None) -> np.ndarray:
    """Marginalise -> read-out error -> big-endian, exactly as QuantumCircuit.py:436-445."""
    qubits = list(qubits)
    if len(qubits) < num_qubits:
        probs = _marginal_little_endian(probs, num_qubits, qubits)
    probs = np.require(probs, dtype=np.float64, requirements=["C", "W", "O"])
    # The marginal is little-endian over the kept qubits in ascending order, so qubit q sits at bit rank(q); its read-out
    # matrix is measurement_error[q].  For "all qubits in order" this is literally the reference's call
    # (QuantumCircuit.py:438-444); for any other subset the reference strides by the qubit id and reads out of bounds
    # (MeasurementErrorApplicator.cpp:109-112) -- here every subset works.
    kept = sorted(qubits)
    ranks = [kept.index(q) for q in qubits]
    mats = {r: m for r, m in zip(ranks, simulator._matrices_for(measurement_error, qubits))}
    (apply_measurement_error or simulator.apply_measurement_error)(probs, mats, ranks, len(qubits), num_cores)
    m = len(qubits)
    return probs.reshape([2] * m).transpose(list(range(m))[::-1]).reshape(-1)


def _validate(num_qubits, instruction_list, qubits, num_trajectories):
    # QuantumCircuit.py:391-409
    if not isinstance(qubits, list) or any(not isinstance(q, int) for q in qubits):
        raise TypeError("Qubits must be a list of integers.")
    if any((q < 0 or q >= num_qubits) for q in qubits):
        raise ValueError(f"One or more qubits are out of range. The valid range is from 0 to {num_qubits - 1}.")
    if instruction_list == []:
        raise ValueError("No instructions in the circuit to execute.")
    if not isinstance(num_trajectories, int):
        raise TypeError("num_trajectories must be an integer.")
    if num_trajectories < 1:
        raise ValueError("num_trajectories must be a positive integer.")


def execute(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise, measurement_error, qubits=None,
            num_trajectories=100, num_cores=1, rng="philox", seed=42, executor=None, traj_offset=0, **engine_options) -> np.ndarray:
    """QuantumCircuit.execute for sim_backend "b200": big-endian probabilities of the measured qubits.

    executor: a RemoteExecutor to reuse (it keeps the compiled program of the last circuit); traj_offset: global id of
    the first trajectory (keys the counter-based RNG)."""
    if qubits is None:
        qubits = list(range(num_qubits))
    _validate(num_qubits, instruction_list, qubits, num_trajectories)
    solver = executor or RemoteExecutor(num_qubits, single_qubit_noise, two_qubit_noise, num_cores, rng=rng, seed=seed, **engine_options)
    # the sum over trajectories stays on the device; marginal, read-out error and bit reversal run there too and only the
    # 2^len(qubits) result comes back
    acc = _accumulator(1 << num_qubits)
    solver.run_sum(instruction_list, int(traj_offset), num_trajectories, device_probs=acc.ptr)
    return simulator.execute_tail_device(acc.ptr, num_qubits, qubits, measurement_error, 1.0 / float(num_trajectories))


_acc_cache = {}


def _accumulator(count: int):
    """Device accumulator float64[count], kept between calls (a cudaMalloc / cudaFree pair per execute() costs up to
    100 ms next to a multi-GB state batch)."""
    buf = _acc_cache.get(count)
    if buf is None or not buf.ptr:
        _acc_cache.clear()
        buf = _acc_cache[count] = simulator.DeviceBuffer(count)
    return buf


def execute_sharded(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise, measurement_error, qubits=None,
                    num_trajectories=100, rng="philox", seed=42, group=None, shard_runner=None,
                    apply_measurement_error=None, executor=None, traj_offset=0, **engine_options) -> np.ndarray:
    """QuantumCircuitMPI.execute over torch.distributed.  Every rank returns the full result (the reference
    broadcasts it, QuantumCircuitMPI.py:367).

    shard_runner(local_start, local_count) -> float64[2^n] sum of |psi|^2 over the shard; default: the GPU engine
    writing into a device tensor that is all-reduced in place (NCCL).  Tests inject a CPU runner (gloo).
    executor: a RemoteExecutor to reuse; traj_offset: global id of the first trajectory of the call.
    """
    import torch
    import torch.distributed as dist

    if qubits is None:
        qubits = list(range(num_qubits))
    _validate(num_qubits, instruction_list, qubits, num_trajectories)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    start, count = shard_range(num_trajectories, rank, world)
    dim = 1 << num_qubits
    if shard_runner is None:
        solver = executor or RemoteExecutor(num_qubits, single_qubit_noise, two_qubit_noise, 1, rng=rng, seed=seed, **engine_options)
        total = torch.zeros(dim, dtype=torch.float64, device="cuda")
        if count > 0:
            solver.run_sum(instruction_list, int(traj_offset) + start, count, device_probs=total.data_ptr(),
                           stream=torch.cuda.current_stream().cuda_stream)
    else:
        local = shard_runner(start, count) if count > 0 else np.zeros(dim)
        total = torch.from_numpy(np.ascontiguousarray(local, np.float64).copy())
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)       # comm.Reduce(SUM) + bcast, :352, :367
    if shard_runner is None and apply_measurement_error is None:
        # GPU run: the tail stays on the device, only the 2^len(qubits) result is copied back
        return simulator.execute_tail_device(total.data_ptr(), num_qubits, qubits, measurement_error, 1.0 / float(num_trajectories),
                                             stream=torch.cuda.current_stream().cuda_stream)
    probs = total.cpu().numpy() / float(num_trajectories)
    return postprocess(probs, num_qubits, qubits, measurement_error, 1, apply_measurement_error)


def install(name: str = "b200", cache_noise_tables: bool = False) -> bool:
    """Registers this package as ``NoisyCircuits.utils.<name>`` and teaches ``QuantumCircuit`` the new
    ``sim_backend``.  Returns False when NoisyCircuits is not importable (nothing is changed then).

    The reference selects its in-process (non-Ray) path by string equality with "custom"
    (QuantumCircuit.py:410), so ``execute`` is wrapped: for the new back-end it runs the same call chain with this
    package's RemoteExecutor (see INTEGRATION.md for the three-line upstream patch that makes the wrapper unnecessary).

    cache_noise_tables: also cache the noise operators ``BuildModel`` builds at construction on disk (content addressed).
    """
    try:
        from NoisyCircuits.QuantumCircuit import QuantumCircuit as QC
    except Exception:
        return False
    pkg = sys.modules[__package__]
    sys.modules[f"NoisyCircuits.utils.{name}"] = pkg
    if cache_noise_tables:
        # QuantumCircuit.__init__ rebuilds the noise operators with joblib on every construction (QuantumCircuit.py:149-157):
        # put the content-addressed on-disk cache in front of the reference's builder (cache.cached_qubit_gate_model)
        from NoisyCircuits.utils.BuildQubitGateModel import BuildModel
        from . import cache
        if not getattr(BuildModel.build_qubit_gate_model, "_ncb200", False):
            build = BuildModel.build_qubit_gate_model

            def build_cached(self):
                return cache.cached_qubit_gate_model(self, build)

            build_cached._ncb200 = True
            BuildModel.build_qubit_gate_model = build_cached
    if name not in QC.available_sim_backends:
        QC.available_sim_backends.append(name)
    if getattr(QC.execute, "_ncb200", False):
        return True
    original = QC.execute

    def execute_with_b200(self, qubits=None, num_trajectories=100, num_cores=-1, **kw):
        if self.sim_backend != name:
            return original(self, qubits, num_trajectories, num_cores, **kw)
        return execute(self.num_qubits, self.instruction_list, self.single_qubit_error, self.two_qubit_error,
                       self.measurement_error, qubits, num_trajectories, 1)

    execute_with_b200._ncb200 = True
    QC.execute = execute_with_b200
    return True
