"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the C ABI (libncb200.so) and is
checked against the CPU oracle (oracle/) on the same seeded inputs, or against the committed golden vectors generated
from the reference's own extensions.

Tolerances (north_star): replay (same per-instruction uniforms) <= 1e-10 state by state; density matrix <= 1e-10 per
probability; free-running MCWF vs density matrix: total variation <= 3/sqrt(N)."""
import os

import numpy as np
import pytest

from noisycircuits_b200 import _lib, circuits
from noisycircuits_b200 import simulator as ncb_simulator
from noisycircuits_b200.program import PackedCircuit, Program
from noisycircuits_b200.solvers import DensityMatrixSolver, PureStateSolver, RemoteExecutor
from oracle import oracle
from conftest import far_tables

pytestmark = pytest.mark.gpu
TOL = 1e-10


def jumpy_uniforms(rng, count, G):
    U = rng.random((count, G))
    U[count // 2:] = 1 - 0.03 * rng.random((count - count // 2, G)) ** 2
    return U


def philox_numpy(seed, traj, instr):
    """Independent numpy restatement of Philox4x32-10 with the engine's (counter, key) convention."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(instr), int(traj) & 0xffffffff, (int(traj) >> 32) & 0xffffffff, 0x6e636232]
    k0, k1 = seed & 0xffffffff, (seed >> 32) & 0xffffffff
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k1) & 0xffffffff, p0 & 0xffffffff]
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    bits = ((c[0] << 21) ^ (c[1] >> 11)) & ((1 << 53) - 1)
    return bits / 9007199254740992.0


def replay(n, instr, sq, tq, U, mask_fix=True, **opts):
    """Seeded replay harness: the same per-instruction uniforms drive the GPU engine and the oracle.  The engine may
    execute commuting instructions in a tile-friendly order (Program.order(), the same circuit); the oracle runs the
    reference algorithm on that order, each uniform staying attached to its instruction."""
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, **opts)
    probs, states, stats = prog.run_mcwf(U.shape[0], uniforms=U, return_states=True)
    order = prog.order()
    assert sorted(order) == list(range(len(instr)))
    if opts.get("strict_order"):
        assert list(order) == list(range(len(instr)))
    O = oracle.Program(n, [instr[g] for g in order], sq, tq)
    worst, acc = 0.0, np.zeros(1 << n)
    for t in range(U.shape[0]):
        ref = O.trajectory(U[t][order], mask_fix=mask_fix)
        worst = max(worst, float(np.abs(states[t] - ref).max()))
        acc += np.abs(ref) ** 2
    assert worst < TOL, (worst, stats, prog.describe().splitlines()[0])
    assert np.abs(probs - acc).max() < TOL * U.shape[0]
    return stats, prog


# ---------------------------------------------------------------------------------------------------------------
# replay harness: trajectories agree state by state
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,opts", [(1, {}), (2, {}), (4, {}), (7, {}), (9, {}), (12, {}), (6, {"fuse": 0}), (8, {"reg_bits": 4})])
def test_replay_resident_heron(heron_noise, n, opts):
    sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, 4, circuits.coupled_pairs(tq, n), seed=n)
    stats, prog = replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(n), 8, len(instr)), **opts)
    assert prog.query(_lib.Q_RESIDENT) == 1
    assert stats["events"] > 0
    if opts.get("fuse") == 0:
        assert prog.query(_lib.Q_NUM_SEGMENTS) == len(instr)       # many segments, one resident pass list


def test_replay_resident_long_program_spans_several_segments(heron_noise):
    """A long small-state program is cut into several segments (shared-memory staging limit of the tile path); the
    resident kernel has to walk all of them."""
    sq, tq, _, _ = heron_noise
    n = 4
    instr = circuits.heron_ansatz(n, 12, circuits.coupled_pairs(tq, n), seed=2)
    stats, prog = replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(5), 6, len(instr)))
    assert prog.query(_lib.Q_RESIDENT) == 1 and prog.query(_lib.Q_NUM_SEGMENTS) > 1


@pytest.mark.parametrize("n,opts", [
    (13, {}), (14, {"reg_bits": 3}), (15, {"tile_bits": 11}), (16, {}), (14, {"fuse": 0}), (16, {"low_bits": 0, "reg_bits": 3}),
    (14, {"strict_order": True}), (16, {"strict_order": True, "tile_bits": 10}), (15, {"reg_bits": 4}),
])
def test_replay_tiled_heron(heron_noise, n, opts):
    sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, 3, circuits.coupled_pairs(tq, n), seed=n)
    stats, prog = replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(n), 6, len(instr)), **opts)
    assert prog.query(_lib.Q_RESIDENT) == 0
    assert stats["events"] > 0 and stats["hard_events"] > 0


@pytest.mark.parametrize("n,opts", [(5, {}), (10, {}), (14, {}), (14, {"reg_bits": 3}), (16, {"tile_bits": 10})])
def test_replay_eagle_dense_gates_and_norm_tracking(eagle_noise, n, opts):
    sq, tq, _, _ = eagle_noise
    instr = circuits.eagle_layered(n, 3, circuits.coupled_pairs(tq, n), seed=5)
    stats, prog = replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(1), 6, len(instr)), **opts)
    assert prog.query(_lib.Q_NORM_PRESERVING) == 0


def test_replay_non_adjacent_pairs(golden, heron_noise):
    _z, meta = golden
    m = meta["far"]
    replay(m["n"], m["instr"], heron_noise[0], far_tables(heron_noise, meta), jumpy_uniforms(np.random.default_rng(2), 6, len(m["instr"])))


def test_replay_full_size_single_layer(heron_noise):
    """One 20-qubit layer (BASELINE configs[4] shape) against the oracle."""
    sq, tq, _, _ = heron_noise
    n = 20
    instr = circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n), seed=42)
    replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(3), 2, len(instr)))


# ---------------------------------------------------------------------------------------------------------------
# the reference's FFI entry points against its golden vectors
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,noise", [("mcwf_heron4", "heron"), ("mcwf_eagle4", "eagle")])
def test_simulate_circuit_and_run_trajectory_match_reference(golden, heron_noise, eagle_noise, case, noise):
    z, meta = golden
    m = meta[case]
    sq, tq, _, _ = heron_noise if noise == "heron" else eagle_noise
    n = m["n"]
    buf = np.zeros(1 << n, np.complex128)
    ncb_simulator.simulate_circuit(m["instr"], buf, sq, tq, n, True, m["T"], False, 2)
    assert np.abs(buf.real - z[case + "_probs"]).max() < TOL and np.abs(buf.imag).max() == 0
    ncb_simulator.simulate_circuit(m["instr"], buf, sq, tq, n, True, m["T"], False, 2)      # accumulates (+=)
    assert np.abs(buf.real - 2 * z[case + "_probs"]).max() < 2 * TOL
    for row, seed in zip(z[case + "_traj"], m["traj_seeds"]):
        buf = np.zeros(1 << n, np.complex128)
        ncb_simulator.run_trajectory(m["instr"], buf, sq, tq, n, seed, 1)
        assert np.abs(buf.real - row).max() < TOL
    ex = RemoteExecutor(n, sq, tq, 2, rng="reference")
    assert np.abs(ex.run(m["T"], m["instr"]) - z[case + "_probs"]).max() < TOL


def test_non_adjacent_pairs_match_patched_reference(golden, heron_noise):
    z, meta = golden
    m = meta["far"]
    n = m["n"]
    tq = far_tables(heron_noise, meta)
    buf = np.zeros(1 << n, np.complex128)
    ncb_simulator.simulate_circuit(m["instr"], buf, heron_noise[0], tq, n, True, m["T"], False, 1)
    assert np.abs(buf.real - z["mcwf_far_probs"]).max() < TOL
    buf = np.zeros(1 << n, np.complex128)
    ncb_simulator.simulate_circuit(m["instr"], buf, {}, {}, n, False, 1, True, 1)
    assert np.abs(buf - z["pure_far_state"]).max() < 1e-12


def test_pure_state_all_gate_kernels(golden):
    z, meta = golden
    m = meta["pure_adjacent"]
    n = m["n"]
    buf = np.zeros(1 << n, np.complex128)
    ncb_simulator.simulate_circuit(m["instr"], buf, {}, {}, n, False, 1, True, 1)
    assert np.abs(buf - z["pure_adjacent_state"]).max() < 1e-12
    buf = np.zeros(1 << n, np.complex128)
    ncb_simulator.simulate_circuit(m["instr"], buf, {}, {}, n, False, 1, False, 1)
    assert np.abs(buf.real - np.abs(z["pure_adjacent_state"]) ** 2).max() < 1e-12
    s = PureStateSolver(n, m["instr"], 1, True).solve()
    assert np.abs(s - z["pure_adjacent_state"]).max() < 1e-12
    p = PureStateSolver(n, m["instr"], 1, False).solve()
    assert p.dtype == np.float64 and abs(p.sum() - 1) < 1e-12


def test_pure_state_tiled(heron_noise):
    """Noise-free: the reordered program is the same unitary (global phases restored), deep circuit, 8 tiles."""
    n = 15
    _sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, 8, circuits.coupled_pairs(tq, n), seed=8)
    ref = oracle.Program(n, instr, noisy=False).pure_state()
    s = PureStateSolver(n, instr, 1, True).solve()
    assert np.abs(s - ref).max() < 1e-12
    s = PureStateSolver(n, instr, 1, True, strict_order=True).solve()
    assert np.abs(s - ref).max() < 1e-12


def test_measurement_error(golden, heron_noise):
    z, meta = golden
    n = meta["meas"]["n"]
    p = z["meas_in"].copy()
    ncb_simulator.apply_measurement_error(p, {q: heron_noise[2][q] for q in range(n)}, list(range(n)), n, 1)
    assert np.abs(p - z["meas_out"]).max() < 1e-15


# ---------------------------------------------------------------------------------------------------------------
# Philox free-running mode
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [5, 13])
def test_philox_draws_match_host_restatement(heron_noise, n):
    """Free-running GPU run == oracle fed with the same counter-based uniforms (resident and tiled paths)."""
    sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, 2, circuits.coupled_pairs(tq, n), seed=21)
    T, seed, begin = 24, 1234567, 1000
    U = np.array([[philox_numpy(seed, begin + t, g) for g in range(len(instr))] for t in range(T)])
    ref = oracle.Program(n, instr, sq, tq).mcwf(T, uniforms=U, mask_fix=True)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF)
    probs, _ = prog.run_mcwf(T, traj_begin=begin, rng=_lib.RNG_PHILOX, seed=seed)
    assert np.abs(probs / T - ref).max() < TOL


def test_sharding_and_batching_do_not_change_the_result(heron_noise):
    sq, tq, _, _ = heron_noise
    n = 14
    instr = circuits.heron_layered(n, 2, circuits.coupled_pairs(tq, n), seed=2)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF)
    full, _ = prog.run_mcwf(96, seed=7)
    a, _ = prog.run_mcwf(40, traj_begin=0, seed=7)
    b, _ = prog.run_mcwf(56, traj_begin=40, seed=7)
    assert np.abs(full - (a + b)).max() < 1e-12
    c, _ = prog.run_mcwf(96, seed=7, batch=17)
    assert np.abs(full - c).max() < 1e-12
    again, _ = prog.run_mcwf(96, seed=7)
    assert np.array_equal(full, again)                 # deterministic
    assert abs(full.sum() - 96) < 1e-9                 # every trajectory is normalised


def test_shared_jump_free_trajectory_resident(heron_noise):
    """Resident path (n <= 12): trajectories whose draws all pick the base branch are one and the same evolution; block 0
    computes it once and the others are counted.  Same result as independent trajectories and as the oracle."""
    sq, tq, _, _ = heron_noise
    n = 8
    instr = circuits.heron_layered(n, 4, circuits.coupled_pairs(tq, n), seed=11)
    pk = PackedCircuit(n, instr, sq, tq)
    shared = Program(pk, _lib.MODE_MCWF)
    indep = Program(pk, _lib.MODE_MCWF, independent_trajectories=True)
    T, seed = 3000, 5
    a, sa = shared.run_mcwf(T, seed=seed)
    b, sb = indep.run_mcwf(T, seed=seed)
    assert np.abs(a - b).max() < 1e-10 * T / 100
    assert sa["events"] == sb["events"] and 0 < sa["events"] < T          # most trajectories are jump free here
    assert abs(a.sum() - T) < 1e-8
    Tq = 40
    U = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(Tq)])
    ref = oracle.Program(n, instr, sq, tq).mcwf(Tq, uniforms=U, mask_fix=True)
    q, _ = shared.run_mcwf(Tq, seed=seed)
    assert np.abs(q / Tq - ref).max() < TOL
    # explicit uniforms: all 0.5 -> every trajectory is the trunk
    flat, st = shared.run_mcwf(16, rng=_lib.RNG_EXPLICIT, uniforms=np.full((16, len(instr)), 0.5))
    one, _ = indep.run_mcwf(1, rng=_lib.RNG_EXPLICIT, uniforms=np.full((1, len(instr)), 0.5))
    assert st["events"] == 0 and np.abs(flat - 16 * one).max() < 1e-12


def test_shared_no_jump_prefix_is_the_same_computation(heron_noise):
    """The batched run computes the evolution all trajectories share before their first jump once per batch
    (ncb_schedule.h).  Against a program compiled with independent_trajectories: same probabilities (only the order of
    the final sum changes), same events, fewer sweeps; and both equal the oracle fed with the same uniforms."""
    sq, tq, _, _ = heron_noise
    n = 14
    instr = circuits.heron_layered(n, 8, circuits.coupled_pairs(tq, n), seed=4)
    pk = PackedCircuit(n, instr, sq, tq)
    shared = Program(pk, _lib.MODE_MCWF)
    indep = Program(pk, _lib.MODE_MCWF, independent_trajectories=True)
    T, seed = 200, 99
    a, sa = shared.run_mcwf(T, seed=seed)
    b, sb = indep.run_mcwf(T, seed=seed)
    assert np.abs(a - b).max() < 1e-12
    assert sa["events"] == sb["events"] and sa["hard_events"] == sb["hard_events"] and sa["events"] > 0
    assert sa["sweeps"] < sb["sweeps"]
    assert abs(a.sum() - T) < 1e-9
    c, _ = shared.run_mcwf(T, seed=seed, batch=33)          # several batches, each with its own trunk
    assert np.abs(a - c).max() < 1e-12
    Tq = 12
    U = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(Tq)])
    ref = oracle.Program(n, instr, sq, tq).mcwf(Tq, uniforms=U, mask_fix=True)
    q, _ = shared.run_mcwf(Tq, seed=seed)
    assert np.abs(q / Tq - ref).max() < TOL


# ---------------------------------------------------------------------------------------------------------------
# density-matrix path and the statistical check
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 4, 6, 7])
def test_density_matrix_vs_restatement(heron_noise, eagle_noise, n):
    for noise, gen in ((heron_noise, circuits.heron_layered), (eagle_noise, circuits.eagle_layered)):
        sq, tq, _, _ = noise
        instr = gen(n, 3, circuits.coupled_pairs(tq, n), seed=9)
        ref = oracle.density_matrix_probs(n, instr, sq, tq)
        got = DensityMatrixSolver(n, sq, tq, instr, 1).solve(list(range(n)))
        assert np.abs(got - ref).max() < TOL
        keep = list(range(0, n, 2))
        assert np.abs(DensityMatrixSolver(n, sq, tq, instr, 1).solve(keep) - oracle.marginal_little_endian(ref, n, keep)).max() < TOL


def test_mcwf_distribution_matches_density_matrix(heron_noise):
    """BASELINE configs[0]: 4-qubit Heron CZ/RZZ ansatz, 1000 trajectories vs the density-matrix method."""
    sq, tq, _, _ = heron_noise
    n, N = 4, 1000
    instr = circuits.heron_ansatz(n, 4, circuits.coupled_pairs(tq, n), seed=42)
    dm = DensityMatrixSolver(n, sq, tq, instr, 1).solve(list(range(n)))
    mc = RemoteExecutor(n, sq, tq, 2).run(N, instr)
    assert abs(mc.sum() - 1) < 1e-9
    assert 0.5 * np.abs(mc - dm).sum() <= 3 / np.sqrt(N)
    hell = np.sqrt(0.5 * np.sum((np.sqrt(mc) - np.sqrt(dm)) ** 2))
    assert hell < 0.05                          # the reference's own acceptance test (tests/test_Integration.py:157-169)


def test_mcwf_distribution_eagle_matches_density_matrix(eagle_noise):
    sq, tq, _, _ = eagle_noise
    n, N = 7, 4000
    instr = circuits.eagle_layered(n, 4, circuits.coupled_pairs(tq, n), seed=3)
    dm = DensityMatrixSolver(n, sq, tq, instr, 1).solve(list(range(n)))
    mc = RemoteExecutor(n, sq, tq, 2).run(N, instr)
    assert 0.5 * np.abs(mc - dm).sum() <= 3 / np.sqrt(N)


def test_mcwf_distribution_tiled_reordered_matches_density_matrix(heron_noise):
    """Free-running, tile path (n = 9 in 2^8-amplitude tiles), reordered execution, noise amplified 40x so that jump
    events (incl. state-dependent ones) are frequent: the distribution still is the density-matrix one."""
    sq, tq, _, _ = heron_noise
    n, N = 9, 6000

    def amplify(ops, f=40.0):
        # p_k -> f p_k for the error branches, identity branch rescaled: stays CPTP for these mixed-unitary + reset lists
        w = [np.trace(k.conj().T @ k).real / k.shape[0] for k in ops]
        err = sum(w[1:])
        out = [np.sqrt(max(1 - f * err, 0.0) / w[0]) * ops[0]] + [np.sqrt(f) * k for k in ops[1:]]
        return out
    good = [q for q in range(n) if q not in (4, 9)]          # T2 <= T1 qubits: every list starts with the scaled identity
    sq2 = {q: {g: (amplify(ops) if q in good else ops) for g, ops in gates.items()} for q, gates in sq.items()}
    instr = circuits.heron_layered(n, 5, circuits.coupled_pairs(tq, n), seed=12)
    dm = DensityMatrixSolver(n, sq2, tq, instr, 1).solve(list(range(n)))
    ex = RemoteExecutor(n, sq2, tq, 2, tile_bits=8)
    mc = ex.run(N, instr)
    assert ex.last_stats["events"] > N // 2 and ex.last_stats["hard_events"] > N // 20 and ex.last_stats["sweeps"] > 0
    assert abs(mc.sum() - 1) < 1e-9
    assert 0.5 * np.abs(mc - dm).sum() <= 3 / np.sqrt(N)


def test_execute_chain_and_full_size_properties(heron_noise):
    from noisycircuits_b200.backend import execute
    sq, tq, meas, _ = heron_noise
    n = 20
    instr = circuits.heron_layered(n, 2, circuits.coupled_pairs(tq, n), seed=42)
    out = execute(n, instr, sq, tq, {q: meas[q] for q in range(n)}, None, 8)
    assert out.shape == (1 << n,) and abs(out.sum() - 1) < 1e-9 and out.min() >= 0
    # fused and unfused schedules are the same function (both strictly in program order)
    a, _ = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, fuse=1, strict_order=True).run_mcwf(4, seed=3)
    b, _ = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, fuse=0).run_mcwf(4, seed=3)
    assert np.abs(a - b).max() < TOL


def test_execute_tail_on_device_matches_host_chain(heron_noise):
    """Marginal -> read-out error -> big-endian on the device (ncb_execute_tail_device) against the host chain of
    QuantumCircuit.py:436-445 (backend.postprocess), and execute() end to end against solver.run + host chain."""
    from noisycircuits_b200 import simulator
    from noisycircuits_b200.backend import execute, postprocess
    from noisycircuits_b200.solvers import RemoteExecutor
    sq, tq, meas, _ = heron_noise
    rng = np.random.default_rng(5)
    n = 9
    p = rng.random(1 << n)
    p /= p.sum()
    import torch                                         # only to put the test vector on the device
    t = torch.from_numpy(p).cuda()
    for qubits in (list(range(n)), [0, 1, 2], [0, 1, 2, 3, 4]):
        m = {q: meas[q] for q in range(len(qubits))}
        dev = simulator.execute_tail_device(t.data_ptr(), n, qubits, m, 0.5)
        ref = postprocess(0.5 * p, n, qubits, m, 1)
        assert dev.shape == ref.shape and np.abs(dev - ref).max() < 1e-14
    no_err = simulator.execute_tail_device(t.data_ptr(), n, [1, 3], None, 1.0)
    full = p.reshape([2] * n)
    drop_axes = tuple(ax for ax in range(n) if (n - 1 - ax) not in (1, 3))
    marg = full.sum(axis=drop_axes).reshape(-1)                 # little-endian over (1, 3)
    assert np.abs(no_err - marg.reshape(2, 2).T.reshape(-1)).max() < 1e-14
    sub = simulator.execute_tail_device(t.data_ptr(), n, [5, 7], {5: meas[5], 7: meas[7]}, 1.0)    # any subset works
    assert np.abs(sub - postprocess(p.copy(), n, [5, 7], {q: meas[q] for q in range(n)}, 1)).max() < 1e-14
    buf = simulator.DeviceBuffer(16)                     # the C ABI's own scratch allocator
    assert buf.ptr != 0
    buf.free()
    n = 10
    instr = circuits.heron_layered(n, 2, circuits.coupled_pairs(tq, n), seed=8)
    mdict = {q: meas[q] for q in range(n)}
    out = execute(n, instr, sq, tq, mdict, None, 64, seed=11)
    ref = postprocess(RemoteExecutor(n, sq, tq, 1, seed=11).run(64, instr), n, list(range(n)), mdict, 1)
    assert np.abs(out - ref).max() < 1e-13


def test_run_many_equals_a_loop_of_run(heron_noise):
    """Parameter sweep over one gate skeleton (QNN training loop): the pipelined run_many returns what run returns."""
    from noisycircuits_b200.solvers import RemoteExecutor
    sq, tq, _, _ = heron_noise
    n = 8
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
    sweeps = [circuits.qnn_circuit(n, 2, pairs, seed=s) for s in range(5)]
    ex = RemoteExecutor(n, sq, tq, 1, seed=3)
    many = ex.run_many(500, sweeps)
    assert len(many) == 5
    for instr, got in zip(sweeps, many):
        ref = RemoteExecutor(n, sq, tq, 1, seed=3).run(500, instr)
        assert np.array_equal(got, ref)
    assert ex.run_many(10, []) == []


@pytest.mark.parametrize("case", range(int(os.environ.get("NCB_FUZZ_CASES", "10"))))
def test_random_circuits_all_gates_random_channels_gpu(case):
    """The CPU suite's differential fuzz (tests/test_planner_emulated.py) on the real kernels: every gate name, random
    channels (Pauli / damping / dense), random pairs, resident and tiled geometries, replayed state by state."""
    from conftest import random_noisy_circuit
    rng = np.random.default_rng(5000 + case)
    n = int(rng.integers(3, 15))
    lean = case % 4 >= 2                          # half of the cases stay on the lean kernels (direct kernel, fused rotations)
    instr, sq, tq = random_noisy_circuit(rng, n, length=int(rng.integers(20, 60)), lean=lean)
    opts = {}
    if n >= 10 and case % 2 == 0:
        opts["tile_bits"] = 9                     # tiled path below the default residency threshold
    if case % 3 == 0 and not lean:
        opts["reg_bits"] = 4
    U = np.random.default_rng(case).random((4, len(instr)))
    U[2:] = 1 - 0.03 * np.random.default_rng(case + 1).random((2, len(instr))) ** 2      # jump-heavy
    stats, _ = replay(n, instr, sq, tq, U, **opts)
    assert stats["events"] > 0
    # free-running batched run (shared prefix, direct kernel, in-line jumps) against the oracle fed the same uniforms
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, **opts)
    T, seed = 6, 77 + case
    order = prog.order()
    Uq = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(T)])
    ref = oracle.Program(n, [instr[g] for g in order], sq, tq).mcwf(T, uniforms=Uq[:, order], mask_fix=True)
    got, _ = prog.run_mcwf(T, seed=seed)
    assert np.abs(got / T - ref).max() < TOL


def test_specialised_kernels_are_the_interpreter_written_out(heron_noise):
    """ncb_options.specialize: one generated kernel per segment (pass programs with literal coefficients) instead of the
    interpreting direct kernel.  Same operators in the same order: the probabilities agree to rounding with the
    interpreter's, the trajectories with the oracle's; the source really is straight-line code."""
    sq, tq, _, _ = heron_noise
    n = 14
    instr = circuits.heron_layered(n, 6, circuits.coupled_pairs(tq, n), seed=12)
    pk = PackedCircuit(n, instr, sq, tq)
    spec = Program(pk, _lib.MODE_MCWF, specialize=True)
    if not spec.specialize_build():
        pytest.skip("nvcc not usable on this box: the interpreter is the only path")
    src = spec.specialized_source(0, 1)
    assert "ncb_seg_0" in src and "spec_rx" in src and "switch" not in src
    plain = Program(pk, _lib.MODE_MCWF)
    T, seed = 300, 21
    a, sa = spec.run_mcwf(T, seed=seed)
    b, sb = plain.run_mcwf(T, seed=seed)
    assert np.abs(a - b).max() < 1e-12 and sa["events"] == sb["events"] and sa["sweeps"] == sb["sweeps"]
    assert sa["specialized_launches"] == sa["tile_launches"] > 0
    assert sb["specialized_launches"] == 0 or os.environ.get("NCB_SPECIALIZE")
    assert abs(a.sum() - T) < 1e-9
    Tq = 10
    order = spec.order()
    U = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(Tq)])
    ref = oracle.Program(n, [instr[g] for g in order], sq, tq).mcwf(Tq, uniforms=U[:, order], mask_fix=True)
    q, _ = spec.run_mcwf(Tq, seed=seed)
    assert np.abs(q / Tq - ref).max() < TOL


# ---------------------------------------------------------------------------------------------------------------
# the benchmarked path at BASELINE sizes, in PROGRAM order as well as in the engine's order
# ---------------------------------------------------------------------------------------------------------------
def _oracle_threads():
    return max(1, min(os.cpu_count() or 1, 16))


def _both_orders(n, instr, sq, tq, prog, U):
    """Oracle trajectories driven by U (indexed like `instr`) in the caller's order and in the engine's order: returns
    (mean |psi|^2 in program order, in engine order, per-trajectory flag: same Kraus branch at every instruction)."""
    order = prog.order()
    P_prog = oracle.Program(n, instr, sq, tq)
    P_eng = oracle.Program(n, [instr[g] for g in order], sq, tq)
    same = []
    for t in range(U.shape[0]) if n <= 16 else []:
        _s, c_prog, _p = P_prog.trajectory(U[t], mask_fix=True, diagnostics=True)
        _s, c_eng, _p = P_eng.trajectory(U[t][order], mask_fix=True, diagnostics=True)
        same.append(bool(np.array_equal(c_prog[order], c_eng)))
    from concurrent.futures import ThreadPoolExecutor          # both orders at once (ctypes releases the GIL)
    Ue = np.ascontiguousarray(U[:, order])
    th = max(1, _oracle_threads() // 2)
    with ThreadPoolExecutor(max_workers=2) as pool:
        fa = pool.submit(P_prog.mcwf, U.shape[0], U, 42, True, th)
        fb = pool.submit(P_eng.mcwf, U.shape[0], Ue, 42, True, th)
        a, b = fa.result(), fb.result()
    return a, b, same


@pytest.mark.parametrize("n,depth", [(13, 6), (16, 4)])
def test_reordered_engine_equals_the_reference_in_program_order(heron_noise, n, depth):
    """The default (tile-friendly) execution order against the oracle run in the CALLER's order.  Operators on disjoint
    qubits commute exactly, so a trajectory differs between the two orders only if a uniform falls between the two
    orders' cumulative branch probabilities (a set of measure ~ noise^2): for every trajectory whose branch sequence is
    the same in both orders -- all of them here -- the final states agree to rounding."""
    sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, depth, circuits.coupled_pairs(tq, n), seed=100 + n)
    T = 8
    U = jumpy_uniforms(np.random.default_rng(n), T, len(instr))
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF)
    order = prog.order()
    assert list(order) != list(range(len(instr)))
    _probs, states, stats = prog.run_mcwf(T, uniforms=U, return_states=True)
    assert stats["hard_events"] > 0
    P_prog = oracle.Program(n, instr, sq, tq)
    P_eng = oracle.Program(n, [instr[g] for g in order], sq, tq)
    agree = 0
    for t in range(T):
        ref, c_prog, _ = P_prog.trajectory(U[t], mask_fix=True, diagnostics=True)
        ref_e, c_eng, _ = P_eng.trajectory(U[t][order], mask_fix=True, diagnostics=True)
        assert np.abs(states[t] - ref_e).max() < TOL                 # always: the reference algorithm on the engine's order
        if np.array_equal(c_prog[order], c_eng):
            agree += 1
            assert np.abs(states[t] - ref).max() < TOL               # and the reference's own trajectory
    assert agree == T
    # strict_order: the caller's order itself
    strict = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, strict_order=True)
    _p, st2, _ = strict.run_mcwf(T, uniforms=U, return_states=True)
    for t in range(T):
        assert np.abs(st2[t] - P_prog.trajectory(U[t], mask_fix=True)).max() < TOL


def test_reference_rng_at_tiled_size_matches_oracle_in_program_order(heron_noise):
    """ncb_simulate_circuit / ncb_run_trajectory / rng="reference" above the tile size (n = 14): compiled in strict program
    order, so the mt19937_64(42 + t) streams meet the same states as in the reference (Simulator.cpp:76-94, 118-133)."""
    sq, tq, _, _ = heron_noise
    n, T = 14, 6
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)        # adjacent pairs: the verbatim reference is right here
    instr = circuits.heron_layered(n, 3, pairs, seed=77)
    O = oracle.Program(n, instr, sq, tq)
    ref = O.mcwf(T, seed0=42, mask_fix=False, threads=_oracle_threads())
    buf = np.zeros(1 << n, np.complex128)
    ncb_simulator.simulate_circuit(instr, buf, sq, tq, n, True, T, False, 2)
    assert np.abs(buf.real - ref).max() < TOL
    one = np.zeros(1 << n, np.complex128)
    ncb_simulator.run_trajectory(instr, one, sq, tq, n, 45, 1)
    assert np.abs(one.real - np.abs(O.trajectory(O.reference_uniforms(45), mask_fix=False)) ** 2).max() < TOL
    assert np.abs(RemoteExecutor(n, sq, tq, 2, rng="reference").run(T, instr) - ref).max() < TOL
    with pytest.raises(_lib.NcbError, match="strict_order"):         # a re-ordered program refuses the reference's streams
        Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF).run_mcwf(2, rng=_lib.RNG_MT19937_REFERENCE)


def test_c5_full_depth_specialised_shared_prefix_vs_oracle(heron_noise):
    """BASELINE configs[4] exactly as bench.py runs it: 20 qubits, depth 50 (1950 instructions), kernels generated per
    segment, shared no-jump prefix, in-line jumps, Philox draws -- against the oracle fed the same counter-based uniforms,
    in the engine's order AND in program order (about 80 s of CPU per oracle trajectory)."""
    sq, tq, _, _ = heron_noise
    n, depth, T, seed, begin = 20, 50, 4, 42, 3
    instr = circuits.heron_layered(n, depth, circuits.coupled_pairs(tq, n), seed=42)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, specialize=True)
    got, stats = prog.run_mcwf(T, traj_begin=begin, seed=seed)
    assert stats["events"] > 0 and abs(got.sum() - T) < 1e-9
    assert stats.get("specialized_launches", 1) > 0
    U = np.array([[philox_numpy(seed, begin + t, g) for g in range(len(instr))] for t in range(T)])
    in_program_order, in_engine_order, _ = _both_orders(n, instr, sq, tq, prog, U)
    assert np.abs(got / T - in_engine_order).max() < TOL
    assert np.abs(got / T - in_program_order).max() < TOL


def test_c4_eagle16_depth40_vs_oracle(eagle_noise):
    """BASELINE configs[3] shape: 16-qubit Eagle ecr/sx/x/rz circuit, depth 40, Philox trajectories against the oracle in
    both orders (non-unitary base operators: the norm is carried and divided out at the end)."""
    sq, tq, _, _ = eagle_noise
    n, depth, T, seed = 16, 40, 3, 9
    instr = circuits.eagle_layered(n, depth, circuits.coupled_pairs(tq, n), seed=42)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF)
    got, stats = prog.run_mcwf(T, seed=seed)
    assert abs(got.sum() - T) < 1e-9
    U = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(T)])
    in_program_order, in_engine_order, same = _both_orders(n, instr, sq, tq, prog, U)
    assert np.abs(got / T - in_engine_order).max() < TOL
    if all(same):
        assert np.abs(got / T - in_program_order).max() < TOL
    else:       # a uniform between the two orders' branch boundaries (measure ~ noise^2): equal in distribution only
        assert 0.5 * np.abs(got / T - in_program_order).sum() < 1.0


def test_c3_qnn12_vs_oracle(heron_noise):
    """BASELINE configs[2] shape: 12-qubit QNN circuit (RY -> x, sx, rz, sx; CZ chain; RX), 4 layers, resident kernel."""
    sq, tq, _, _ = heron_noise
    n = 12
    instr = circuits.qnn_circuit(n, 4, circuits.coupled_pairs(tq, n, adjacent_only=True), seed=42)
    replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(12), 6, len(instr)))
    T, seed = 64, 5
    U = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(T)])
    ref = oracle.Program(n, instr, sq, tq).mcwf(T, uniforms=U, mask_fix=True, threads=_oracle_threads())
    got = RemoteExecutor(n, sq, tq, 1, seed=seed).run(T, instr)
    assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("n,depth", [(10, 10), (11, 2)])
def test_c2_density_matrix_layered_heron_vs_restatement(heron_noise, n, depth):
    """BASELINE configs[1] shape (layered Heron circuit, 4^n vectorised state) at the largest sizes the numpy restatement
    finishes in about a minute; 1e-10 per probability (north star check #1).  Parity unpinned at the qiskit-aer boundary."""
    sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, depth, circuits.coupled_pairs(tq, n), seed=42)
    ref = oracle.density_matrix_probs(n, instr, sq, tq)
    got = DensityMatrixSolver(n, sq, tq, instr, 1).solve(list(range(n)))
    assert abs(got.sum() - 1) < 1e-10 and np.abs(got - ref).max() < TOL


def test_c2_density_matrix_12_qubits_properties_and_mcwf_tv(heron_noise):
    """configs[1] at its full size (12 qubits, 4^12 state = 268 MB) through size-independent properties: trace 1,
    non-negative diagonal, and the free-running MCWF distribution of the same circuit within TV <= 3/sqrt(N) of it
    (north star check #3; N = 10^5 trajectories on the resident kernel)."""
    sq, tq, _, _ = heron_noise
    n, N = 12, 100000
    instr = circuits.heron_layered(n, 10, circuits.coupled_pairs(tq, n), seed=42)
    dm = DensityMatrixSolver(n, sq, tq, instr, 1).solve(list(range(n)))
    assert abs(dm.sum() - 1) < 1e-10 and dm.min() > -1e-14
    mc = RemoteExecutor(n, sq, tq, 1).run(N, instr)
    assert 0.5 * np.abs(mc - dm).sum() <= 3 / np.sqrt(N)


def test_partial_measurement_on_device_any_subset(heron_noise):
    """ncb_execute_tail_device on measured subsets that are not permutations of 0..m-1 against the dense host chain."""
    from noisycircuits_b200 import simulator
    from noisycircuits_b200.backend import postprocess
    _sq, _tq, meas, _ = heron_noise
    import torch
    n = 6
    p = np.random.default_rng(8).random(1 << n)
    p /= p.sum()
    t = torch.from_numpy(p).cuda()
    mdict = {q: meas[q] for q in range(n)}
    for qubits in ([2, 3], [5], [0, 2, 4], [4, 1], list(range(n))):
        dev = simulator.execute_tail_device(t.data_ptr(), n, qubits, mdict, 1.0)
        ref = postprocess(p.copy(), n, qubits, mdict, 1)
        assert np.abs(dev - ref).max() < 1e-14, qubits


def test_execute_sharded_gpu_branch(heron_noise):
    """backend.execute_sharded on a device (QuantumCircuitMPI.execute: shard -> device accumulator -> all-reduce -> tail on
    the device); with one rank (no process group) it must return what execute() returns, and a trajectory offset must be
    the same as sharding by hand.  The multi-rank NCCL path is the N > 1 end-to-end arm of bench.py."""
    from noisycircuits_b200.backend import execute, execute_sharded
    sq, tq, meas, _ = heron_noise
    n = 13
    instr = circuits.heron_layered(n, 2, circuits.coupled_pairs(tq, n), seed=6)
    mdict = {q: meas[q] for q in range(n)}
    a = execute(n, instr, sq, tq, mdict, None, 96, seed=5)
    b = execute_sharded(n, instr, sq, tq, mdict, None, 96, seed=5)
    assert np.abs(a - b).max() < 1e-14 and abs(b.sum() - 1) < 1e-9
    ex = RemoteExecutor(n, sq, tq, 1, seed=5)
    c = execute_sharded(n, instr, sq, tq, mdict, [0, 5, 7], 40, executor=ex, traj_offset=56)
    d = execute(n, instr, sq, tq, mdict, [0, 5, 7], 40, executor=ex, traj_offset=56)
    assert np.abs(c - d).max() < 1e-14 and c.shape == (8,)


@pytest.mark.parametrize("n", [7, 14])
def test_unitaries_on_five_to_eight_qubits_gpu(heron_noise, n):
    """`unitary` on 5..8 qubits (QuantumGates.hpp:968-1064 takes any k): generic_unitary_kernel segments between ordinary
    ones -- pure state against the oracle (1e-12), noisy replay state by state, density matrix at 5 qubits."""
    sq, tq, _, _ = heron_noise
    rng = np.random.default_rng(40 + n)

    def haar(k):
        d = 1 << k
        return np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0]
    layer = circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n, adjacent_only=True), seed=n)
    instr = []
    for k in (5, 6, min(8, n), 7):
        instr += [["unitary", [int(q) for q in rng.choice(n, size=k, replace=False)], haar(k)]] + layer
    instr += [["unitary", [int(q) for q in rng.choice(n, size=5, replace=False)], haar(5)]]
    ref = oracle.Program(n, instr, noisy=False).pure_state()
    assert np.abs(PureStateSolver(n, instr, 1, True).solve() - ref).max() < 1e-12
    assert np.abs(PureStateSolver(n, instr, 1, True, strict_order=True).solve() - ref).max() < 1e-12
    stats, prog = replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(n), 4, len(instr)))
    assert prog.query(_lib.Q_RESIDENT) == 0 and stats["events"] > 0
    got, _ = prog.run_mcwf(32, seed=3)                              # free running, shared prefix
    assert abs(got.sum() - 32) < 1e-9
    if n == 7:
        m = 5
        dinstr = circuits.heron_layered(m, 1, circuits.coupled_pairs(tq, m), seed=1) + [["unitary", [3, 0, 4, 1, 2], haar(5)]] + \
            circuits.heron_layered(m, 1, circuits.coupled_pairs(tq, m), seed=2)
        dm = DensityMatrixSolver(m, sq, tq, dinstr, 1).solve(list(range(m)))
        assert np.abs(dm - oracle.density_matrix_probs(m, dinstr, sq, tq)).max() < TOL


def test_generated_event_kernels_single_inline_jump(heron_noise):
    """Generated event kernels, the WF_PULL path: a trajectory whose only event in a segment is a jump known from the
    uniform alone keeps the merged operator groups that do not touch the jump's qubits whole around it.  One certain jump
    per trajectory at 48 different instructions (and a few trajectories with a state-dependent one), replayed state by
    state against the oracle; the same with the pull switched off gives the same states."""
    sq, tq, _, _ = heron_noise
    n = 15
    instr = circuits.heron_layered(n, 5, circuits.coupled_pairs(tq, n), seed=31)
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, specialize=True)
    order = prog.order()
    G = len(instr)
    rng = np.random.default_rng(4)
    rows, kinds = [], []
    for g in rng.permutation(G):
        if len(rows) >= 56:
            break
        for trial in range(40):
            u = np.full(G, 0.5)
            u[g] = 1.0 - 10.0 ** (-rng.uniform(0.5, 5.0))
            ev = prog.events(u)
            if len(ev) == 1 and int(ev[0][0]) == g:
                certain = int(ev[0][1]) == int(ev[0][2])
                if certain and kinds.count("certain") < 48 or (not certain and kinds.count("hard") < 8):
                    rows.append(u); kinds.append("certain" if certain else "hard")
                    break
    assert kinds.count("certain") >= 30 and kinds.count("hard") >= 2
    U = np.array(rows)
    O = oracle.Program(n, [instr[g] for g in order], sq, tq)
    _p, states, stats = prog.run_mcwf(U.shape[0], uniforms=U, return_states=True)
    assert stats["events"] == U.shape[0] and stats["specialized_launches"] == stats["tile_launches"] > 0
    worst = max(float(np.abs(states[t] - O.trajectory(U[t][order], mask_fix=True)).max()) for t in range(U.shape[0]))
    assert worst < TOL, worst
    # batched (shared prefix: every trajectory forks from the trunk at the segment of its jump)
    probs, _ = prog.run_mcwf(U.shape[0], rng=_lib.RNG_EXPLICIT, uniforms=U)
    ref = sum(np.abs(O.trajectory(U[t][order], mask_fix=True)) ** 2 for t in range(U.shape[0]))
    assert np.abs(probs - ref).max() < TOL * U.shape[0]


# ---------------------------------------------------------------------------------------------------------------
# generated resident kernels (states that fit one CTA: trunk with checkpoints, trajectories resumed from them)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,opts", [(5, {}), (9, {}), (10, {"reg_bits": 4}), (11, {}), (12, {}), (12, {"reg_bits": 3}),
                                    (12, {"strict_order": True})])
def test_replay_generated_resident_kernels(heron_noise, n, opts):
    """The replay harness on the generated resident kernels (ncb_res_trunk / ncb_res): every trajectory starts from the trunk's
    checkpoint before its first flagged draw and has to end in the oracle's state."""
    sq, tq, _, _ = heron_noise
    instr = circuits.heron_layered(n, 4, circuits.coupled_pairs(tq, n), seed=n)
    U = jumpy_uniforms(np.random.default_rng(n), 10, len(instr))
    U[0] = 0.5                                   # a trajectory that never leaves the trunk
    stats, prog = replay(n, instr, sq, tq, U, specialize="parametric", **opts)
    assert prog.query(_lib.Q_RESIDENT) == 1
    assert stats["specialized_launches"] == 2 and stats["events"] > 0 and stats["hard_events"] > 0


def test_replay_generated_resident_kernels_eagle(eagle_noise):
    """Eagle at a resident size: generic 2x2 factors of the decomposed ecr, norm tracking."""
    sq, tq, _, _ = eagle_noise
    n = 10
    instr = circuits.eagle_layered(n, 3, circuits.coupled_pairs(tq, n), seed=5)
    stats, prog = replay(n, instr, sq, tq, jumpy_uniforms(np.random.default_rng(1), 8, len(instr)), specialize="parametric")
    assert stats["specialized_launches"] == 2 and prog.query(_lib.Q_NORM_PRESERVING) == 0


@pytest.mark.parametrize("n,layers", [(12, 4), (7, 3)])
def test_generated_resident_kernels_equal_the_interpreter(heron_noise, n, layers):
    """Probability path (no states): Philox draws, trajectories that stay on the trunk are only counted; the same sum as the
    interpreting resident kernel and as the oracle, and independent of how the call is split."""
    sq, tq, _, _ = heron_noise
    instr = circuits.qnn_circuit(n, layers, circuits.coupled_pairs(tq, n, adjacent_only=True), seed=7)
    pk = PackedCircuit(n, instr, sq, tq)
    gen = Program(pk, _lib.MODE_MCWF, specialize="parametric")
    interp = Program(pk, _lib.MODE_MCWF)
    T, seed = 3000, 9
    a, sa = gen.run_mcwf(T, seed=seed)
    b, sb = interp.run_mcwf(T, seed=seed)
    assert sa["specialized_launches"] == 2 and sb["specialized_launches"] == 0
    assert sa["events"] == sb["events"] and sa["hard_events"] == sb["hard_events"] and sa["events"] > 0
    assert np.abs(a - b).max() < 1e-12 * T and abs(a.sum() - T) < 1e-8
    c1, _ = gen.run_mcwf(1000, traj_begin=0, seed=seed)
    c2, _ = gen.run_mcwf(2000, traj_begin=1000, seed=seed)
    assert np.abs(c1 + c2 - a).max() < 1e-12 * T
    Tq = 48
    U = np.array([[philox_numpy(seed, t, g) for g in range(len(instr))] for t in range(Tq)])
    ref = oracle.Program(n, instr, sq, tq).mcwf(Tq, uniforms=U, mask_fix=True, threads=_oracle_threads())
    q, _ = gen.run_mcwf(Tq, seed=seed)
    assert np.abs(q / Tq - ref).max() < TOL
    # an in-place update (parameter sweep) keeps the loaded kernels and uploads the new tables
    instr2 = circuits.qnn_circuit(n, layers, circuits.coupled_pairs(tq, n, adjacent_only=True), seed=8)
    gen.update(PackedCircuit(n, instr2, sq, tq))
    d, sd = gen.run_mcwf(T, seed=seed)
    e, _ = Program(PackedCircuit(n, instr2, sq, tq), _lib.MODE_MCWF).run_mcwf(T, seed=seed)
    assert sd["specialized_launches"] == 2 and np.abs(d - e).max() < 1e-12 * T


def test_generated_resident_kernels_are_deterministic_across_updates(heron_noise):
    """Two circuits taking turns in ONE engine object (the parameter-sweep path: tables re-uploaded, trunk re-run beside the
    planning kernels): every repeat returns the same bits.  (A missing barrier after the trunk's checkpoint copy once let a
    fast warp's results into a checkpoint in 2 % of the runs.)"""
    sq, tq, _, _ = heron_noise
    n, T = 12, 4000
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
    pk = [PackedCircuit(n, circuits.qnn_circuit(n, 4, pairs, seed=100 + k), sq, tq) for k in range(2)]
    prog = Program(pk[0], _lib.MODE_MCWF, specialize="parametric")
    refs = []
    for k in range(2):
        prog.update(pk[k])
        refs.append(prog.run_mcwf(T)[0])
    fresh = Program(pk[1], _lib.MODE_MCWF, specialize="parametric").run_mcwf(T)[0]
    assert np.array_equal(fresh, refs[1])
    for r in range(120):
        prog.update(pk[r & 1])
        p, st = prog.run_mcwf(T)
        assert st["specialized_launches"] == 2 and np.array_equal(p, refs[r & 1]), r


def test_two_phase_run_keeps_several_programs_in_flight(heron_noise):
    """ncb_mcwf_launch / ncb_mcwf_collect: two resident programs launched back to back and collected afterwards return what
    ncb_mcwf_run returns, bit for bit; a tiled program goes through the same calls (it runs to completion in the launch); a
    second launch on a program that has one in flight is refused."""
    sq, tq, _, _ = heron_noise
    n, T = 11, 3000
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
    progs = [Program(PackedCircuit(n, circuits.qnn_circuit(n, 3, pairs, seed=40 + k), sq, tq), _lib.MODE_MCWF, specialize="parametric") for k in range(2)]
    ref = [p.run_mcwf(T, seed=3)[0] for p in progs]
    handles = [p.launch_mcwf(T, seed=3) for p in progs]
    with pytest.raises(_lib.NcbError):
        progs[0].launch_mcwf(T, seed=3)
    for p, h, r in zip(progs, handles, ref):
        got, st = p.collect_mcwf(h)
        assert np.array_equal(got, r) and st["specialized_launches"] == 2 and st["events"] > 0
    with pytest.raises(_lib.NcbError):
        progs[0].collect_mcwf(handles[0])
    nt = 14
    tiled = Program(PackedCircuit(nt, circuits.heron_layered(nt, 2, circuits.coupled_pairs(tq, nt), seed=4), sq, tq), _lib.MODE_MCWF)
    a, sa = tiled.run_mcwf(64, seed=1)
    b, sb = tiled.collect_mcwf(tiled.launch_mcwf(64, seed=1))
    assert np.array_equal(a, b) and sa["events"] == sb["events"]
