"""The N > 1 path (QuantumCircuitMPI.execute shape: contiguous trajectory shards + one sum-reduction) on CPU with
world_size = 2 and the gloo backend.  The per-shard compute is injected (the CPU oracle stands in for the GPU engine,
which needs a device); what is tested is the host logic: shard ranges, global trajectory ids keying the RNG, the
collective, and the post-processing chain."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.backend import execute_sharded, postprocess
from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    sq, tq, meas, _ = noise_io.load_noise_tables(os.path.join(GOLDEN, "heron20_noise.npz"))
    n = 4
    instr = circuits.heron_ansatz(n, 2, circuits.coupled_pairs(tq, n), seed=3)
    return n, instr, sq, tq, {q: meas[q] for q in range(n)}


def _oracle_meas(probs, measurement_noise, qubit_list, num_qubits, num_threads):
    oracle.measurement_error(probs, measurement_noise, qubit_list, num_qubits)


def _worker(rank, world, port, T, qubits, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, instr, sq, tq, meas = _case()
        P = oracle.Program(n, instr, sq, tq)

        def runner(start, count):
            # reference seeding with GLOBAL trajectory ids: seed = 42 + local_start + i (QuantumCircuitMPI.py:339-350)
            U = np.stack([P.reference_uniforms(42 + start + i) for i in range(count)])
            return P.mcwf(count, uniforms=U, mask_fix=False) * count
        res = execute_sharded(n, instr, sq, tq, meas, qubits, T, shard_runner=runner, apply_measurement_error=_oracle_meas)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T,qubits", [(9, None), (8, [0, 1, 2])])
def test_two_rank_execute_equals_single_process(T, qubits):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), T, qubits, out), nprocs=world, join=True)
    n, instr, sq, tq, meas = _case()
    P = oracle.Program(n, instr, sq, tq)
    full = P.mcwf(T, seed0=42, mask_fix=False)
    ref = postprocess(full, n, qubits if qubits is not None else list(range(n)), meas, 1, _oracle_meas)
    assert np.abs(out[0] - ref).max() < 1e-12
    assert np.array_equal(out[0], out[1])             # every rank holds the reduced result
    assert abs(out[0].sum() - (1.0 if qubits is None else out[0].sum())) < 1e-9
