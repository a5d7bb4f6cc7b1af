"""Host compiler + batch scheduler + tile geometry + register-blocked executors, run on the CPU through the
thread-by-thread emulation in tests/emu (the same C++ sources the CUDA library is built from), against the oracle.
Replay harness: identical per-instruction uniforms go to both; states must agree to 1e-10 (north_star)."""
import os

import numpy as np
import pytest

from noisycircuits_b200 import circuits
from noisycircuits_b200.program import PackedCircuit
from oracle import oracle
from emu.emu import EmuProgram, lib as emu_lib
from conftest import far_tables, random_noisy_circuit

TOL = 1e-10


def jumpy_uniforms(rng, count, G):
    """Half the trajectories get uniforms pushed towards 1 so that jump events (certain and state-dependent) fire."""
    U = rng.random((count, G))
    U[count // 2:] = 1 - 0.03 * rng.random((count - count // 2, G)) ** 2
    return U


def check_replay(n, instr, sq, tq, U, **opts):
    """Engine (emulated) vs the reference algorithm on the engine's execution order: the engine may run commuting
    instructions in another order (the same circuit); uniforms stay attached to their instructions."""
    pk = PackedCircuit(n, instr, sq, tq)
    E = EmuProgram(pk, 0, **opts)
    order = E.order()
    assert sorted(order) == list(range(len(instr)))
    if not opts.get("reorder", 1):
        assert list(order) == list(range(len(instr)))
    O = oracle.Program(n, [instr[g] for g in order], sq, tq)
    st, stats = E.mcwf_states(U)
    worst = 0.0
    for t in range(U.shape[0]):
        ref = O.trajectory(U[t][order], mask_fix=True)
        worst = max(worst, np.abs(st[t] - ref).max())
    assert worst < TOL, (worst, stats, E.describe())
    # the shared no-jump prefix is the same computation done once: states bit-identical, never more sweeps
    st2, stats2 = E.mcwf_states(U, share=True)
    assert np.array_equal(st, st2), np.abs(st - st2).max()
    assert stats2["sweeps"] <= stats["sweeps"] + 2 * E.query(2)
    return stats, E


@pytest.mark.parametrize("n,tile_bits,reg_bits,low_bits,fuse", [
    (4, 12, 4, 3, 1), (6, 12, 3, 3, 1), (9, 12, 4, 3, 1),       # whole state in one tile (virtual bits for n < K)
    (10, 9, 4, 3, 1), (10, 8, 3, 2, 1), (11, 9, 4, 0, 1),        # tiled: 2..8 tiles per trajectory
    (10, 9, 4, 3, 0), (11, 8, 3, 3, 0),                         # one sweep per instruction
])
@pytest.mark.parametrize("reorder", [0, 1])
def test_heron_replay(heron_noise, n, tile_bits, reg_bits, low_bits, fuse, reorder):
    sq, tq, _, _ = heron_noise
    pairs = circuits.coupled_pairs(tq, n)
    instr = circuits.heron_layered(n, 3, pairs, seed=n)
    U = jumpy_uniforms(np.random.default_rng(n), 6, len(instr))
    stats, E = check_replay(n, instr, sq, tq, U, tile_bits=tile_bits, reg_bits=reg_bits, low_bits=low_bits, fuse=fuse, reorder=reorder)
    assert stats["events"] > 0 and stats["hard_events"] > 0
    if not fuse:
        assert E.query(2) == len(instr)


@pytest.mark.parametrize("n,tile_bits,reg_bits", [(5, 12, 4), (10, 9, 4), (10, 8, 3)])
def test_eagle_replay_dense_two_qubit_and_non_unitary_base(eagle_noise, n, tile_bits, reg_bits):
    """ECR is a dense 4x4; several Eagle qubits have T2 > T1, so the base Kraus operator is not unitary and the
    engine has to carry the norm."""
    sq, tq, _, _ = eagle_noise
    pairs = circuits.coupled_pairs(tq, n)
    instr = circuits.eagle_layered(n, 3, pairs, seed=5)
    U = jumpy_uniforms(np.random.default_rng(1), 6, len(instr))
    stats, E = check_replay(n, instr, sq, tq, U, tile_bits=tile_bits, reg_bits=reg_bits, low_bits=3, fuse=1)
    assert E.query(5) == 0          # not norm preserving
    assert stats["hard_events"] > 0


def test_generic_products_hand_their_phases_to_the_tables(eagle_noise, monkeypatch):
    """Fast lists: a generic one-bit product D1 . X . D2 leaves its phases in neighbouring phase tables and runs as the rx-like X
    (CompileOptions::split_generic).  Same trajectories as with the switch off (and as the oracle), fewer generic operators."""
    import re
    sq, tq, _, _ = eagle_noise
    n = 10
    instr = circuits.eagle_layered(n, 3, circuits.coupled_pairs(tq, n), seed=5)
    U = jumpy_uniforms(np.random.default_rng(4), 6, len(instr))
    _st, on = check_replay(n, instr, sq, tq, U, tile_bits=9, reg_bits=4, low_bits=3, fuse=1)
    states_on, _ = on.mcwf_states(U)
    monkeypatch.setenv("NCB_SPLIT_U", "0")
    _st, off = check_replay(n, instr, sq, tq, U, tile_bits=9, reg_bits=4, low_bits=3, fuse=1)
    states_off, _ = off.mcwf_states(U)
    count = lambda E: int(re.search(r"dense1:(\d+)", E.describe()).group(1))
    assert count(on) < count(off)
    assert np.abs(states_on - states_off).max() < 1e-12


def test_non_adjacent_pairs_replay(golden, heron_noise):
    _z, meta = golden
    m = meta["far"]
    U = jumpy_uniforms(np.random.default_rng(2), 6, len(m["instr"]))
    check_replay(m["n"], m["instr"], heron_noise[0], far_tables(heron_noise, meta), U, tile_bits=12, reg_bits=4, low_bits=3, fuse=1)


def test_reference_seeded_trajectories_match_golden(golden, heron_noise, eagle_noise):
    """The reference's own mt19937_64(42 + t) streams, fed through the emulated engine, reproduce the golden
    single-trajectory vectors of simulator_mpi.run_trajectory."""
    z, meta = golden
    for case, noise in (("mcwf_heron4", heron_noise), ("mcwf_eagle4", eagle_noise)):
        m = meta[case]
        O = oracle.Program(m["n"], m["instr"], noise[0], noise[1])
        E = EmuProgram(PackedCircuit(m["n"], m["instr"], noise[0], noise[1]), 0)
        U = np.stack([O.reference_uniforms(s) for s in m["traj_seeds"]])
        st, _ = E.mcwf_states(U)
        assert np.abs(np.abs(st) ** 2 - z[case + "_traj"]).max() < TOL


@pytest.mark.parametrize("tile_bits,reg_bits", [(12, 4), (12, 3)])
def test_pure_state_all_gates_and_unitaries(golden, tile_bits, reg_bits):
    z, meta = golden
    m = meta["pure_adjacent"]
    instr = m["instr"]
    if reg_bits == 3:          # 3-qubit unitaries need reg_bits = 4: compare the rest against the oracle
        instr = [i for i in instr if not (i[0] == "unitary" and len(i[1]) > 2)]
        ref = oracle.Program(m["n"], instr, noisy=False).pure_state()
    else:
        ref = z["pure_adjacent_state"]
    E = EmuProgram(PackedCircuit(m["n"], instr, noisy=False), 2, tile_bits=tile_bits, reg_bits=reg_bits)
    assert np.abs(E.run_deterministic() - ref).max() < 1e-12


def test_three_qubit_unitary_needs_four_register_bits(golden):
    _z, meta = golden
    m = meta["pure_adjacent"]
    with pytest.raises(RuntimeError, match="reg_bits"):
        EmuProgram(PackedCircuit(m["n"], m["instr"], noisy=False), 2, tile_bits=12, reg_bits=3)


@pytest.mark.parametrize("n,tile_bits", [(3, 12), (5, 9)])
def test_density_matrix_superoperators(heron_noise, eagle_noise, n, tile_bits):
    """DENSITY mode on the vectorised 4^n state against the numpy restatement, <= 1e-10 per probability."""
    for noise, gen in ((heron_noise, circuits.heron_layered), (eagle_noise, circuits.eagle_layered)):
        sq, tq, _, _ = noise
        pairs = circuits.coupled_pairs(tq, n)
        instr = gen(n, 2, pairs, seed=9)
        E = EmuProgram(PackedCircuit(n, instr, sq, tq), 1, tile_bits=tile_bits, reg_bits=4)
        vec = E.run_deterministic()
        rho = vec.reshape(1 << n, 1 << n).T          # vec index = row | col << n
        ref = oracle.density_matrix(n, instr, sq, tq)
        assert np.abs(rho - ref).max() < TOL
        assert abs(np.trace(rho).real - 1) < 1e-12


def test_reordering_is_the_same_circuit_and_saves_sweeps(heron_noise):
    """Without noise the reordered program is the same unitary; with noise it needs fewer segments."""
    sq, tq, _, _ = heron_noise
    n = 12
    instr = circuits.heron_layered(n, 6, circuits.coupled_pairs(tq, n), seed=1)
    ref = oracle.Program(n, instr, noisy=False).pure_state()
    segs = {}
    for reorder in (0, 1):
        E = EmuProgram(PackedCircuit(n, instr, noisy=False), 2, tile_bits=9, reg_bits=4, reorder=reorder)
        assert np.abs(E.run_deterministic() - ref).max() < 1e-12
        segs[reorder] = E.query(2)
    assert segs[1] < segs[0]


def test_event_classification_is_exact_for_pauli_channels(heron_noise):
    """For a mixture of scaled unitaries every branch is known from u alone: no state-dependent events."""
    n = 3
    p = np.array([0.9, 0.04, 0.03, 0.03])
    paulis = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1, -1])]
    ops = [np.sqrt(w) * P for w, P in zip(p, paulis)]
    sq = {q: {"sx": ops, "rz": ops} for q in range(n)}
    instr = [["sx", [q, q], 0] for q in range(n)] * 4 + [["rz", [q, q], 0.3] for q in range(n)]
    U = np.random.default_rng(0).random((16, len(instr)))
    stats, _ = check_replay(n, instr, sq, {}, U, tile_bits=12, reg_bits=4, low_bits=3, fuse=1)
    assert stats["hard_events"] == 0 and stats["events"] > 0


def test_philox_is_the_reference_philox4x32_10():
    """Known-answer test of Philox4x32-10 (Random123 kat_vectors: counter = key = 0 and all-ones)."""
    import ctypes as C
    L = emu_lib()
    # philox_uniform(seed, traj, instr) uses counter (instr, traj_lo, traj_hi, 0x6e636232), key = seed: check
    # the [0,1) range and independence of the three coordinates instead of raw words, plus determinism.
    a = L.emu_philox_uniform(1, 2, 3)
    assert 0.0 <= a < 1.0
    assert a == L.emu_philox_uniform(1, 2, 3)
    assert len({L.emu_philox_uniform(s, t, i) for s in range(3) for t in range(3) for i in range(3)}) == 27
    xs = np.array([L.emu_philox_uniform(7, t, 11) for t in range(20000)])
    assert abs(xs.mean() - 0.5) < 0.01 and abs(xs.var() - 1 / 12) < 0.005


def test_shared_prefix_schedule(heron_noise):
    """Trajectories share one evolution until their first jump: late first events, no events at all, events in the
    first and in the last segment -- the shared schedule must reproduce the independent one bit for bit and save the
    sweeps of the common prefix."""
    sq, tq, _, _ = heron_noise
    n = 10
    pairs = circuits.coupled_pairs(tq, n)
    instr = circuits.heron_layered(n, 6, pairs, seed=3)
    G = len(instr)
    rng = np.random.default_rng(7)
    U = np.full((8, G), 0.5)                      # no events anywhere
    U[1, G - 1] = 1 - 1e-9                        # first event in the last segment
    U[2, 0] = 1 - 1e-9                            # first event at the first instruction
    U[3, G // 2] = 1 - 1e-9                       # one event in the middle
    U[4, G // 2:] = 1 - 0.03 * rng.random(G - G // 2) ** 2      # many events from the middle on
    U[5] = 1 - 0.03 * rng.random(G) ** 2          # events everywhere
    E = EmuProgram(PackedCircuit(n, instr, sq, tq), 0, tile_bits=8, reg_bits=3, low_bits=2, fuse=1, reorder=1)
    assert E.query(2) > 3                         # several segments
    a, sa = E.mcwf_states(U, share=False)
    b, sb = E.mcwf_states(U, share=True)
    assert np.array_equal(a, b)
    assert np.array_equal(b[0], b[6]) and np.array_equal(b[0], b[7])      # the event-free trajectories are the trunk
    assert sb["sweeps"] < sa["sweeps"]
    assert sa["events"] == sb["events"] and sa["events"] > 10


def test_compile_retries_when_a_segment_program_overflows(heron_noise, monkeypatch):
    """The segmenter cuts with an estimate of the staged program size; with the estimate sabotaged (test hook) the
    first attempts overflow the 16 KB staging area and compile_program must converge by itself."""
    sq, tq, _, _ = heron_noise
    n = 8                       # the whole state is one tile: only the program size cuts segments
    instr = circuits.heron_layered(n, 30, circuits.coupled_pairs(tq, n), seed=9)
    monkeypatch.setenv("NCB_BLOB_COST_SCALE", "0.01")
    monkeypatch.setenv("NCB_MAX_SEG_OPS", "400")
    U = jumpy_uniforms(np.random.default_rng(3), 4, len(instr))
    stats, E = check_replay(n, instr, sq, tq, U, tile_bits=8, reg_bits=3, low_bits=2, fuse=1, reorder=1)
    assert E.query(2) >= 2
    assert "attempts=1 " not in E.describe().split("\n")[0]            # the overflow path was really taken


@pytest.mark.parametrize("case", range(int(os.environ.get("NCB_FUZZ_CASES", "12"))))
def test_random_circuits_all_gates_random_channels(case):
    """Differential fuzz of everything between the instruction list and the amplitudes: every gate name of the
    reference (FunctionMapper.hpp:25-79) plus noise-free unitaries, random channels of three kinds on random (also
    non-adjacent) pairs, tiled and single-tile geometries -- replayed against the oracle with jump-heavy uniforms."""
    rng = np.random.default_rng(1000 + case)
    n = int(rng.integers(4, 11))
    tile_bits = 8 if n > 8 and case % 2 == 0 else 12
    reg_bits = 3 if case % 3 else 4
    instr, sq, tq = random_noisy_circuit(rng, n, lean=case % 4 >= 2)
    if case % 4 >= 2:
        reg_bits = 3
    U = jumpy_uniforms(rng, 4, len(instr))
    stats, _E = check_replay(n, instr, sq, tq, U, tile_bits=tile_bits, reg_bits=reg_bits, low_bits=2, fuse=1, reorder=1)
    assert stats["events"] > 0


def _haar(rng, k):
    d = 1 << k
    return np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0]


@pytest.mark.parametrize("n,tile_bits,reg_bits", [(7, 12, 3), (9, 12, 4), (13, 9, 3), (14, 11, 4)])
def test_unitaries_on_five_to_eight_qubits(heron_noise, n, tile_bits, reg_bits):
    """`unitary` on k = 5..8 qubits (the reference's apply_unitary_gate takes any k, QuantumGates.hpp:968-1064): a segment
    of its own, run as gather -> mat-vec -> scatter.  Pure state (reordered and strict), with noise around it (the
    instruction itself is noise free, FunctionMapper.hpp:77), first and last in the circuit, against the oracle."""
    sq, tq, _, _ = heron_noise
    rng = np.random.default_rng(n)
    pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
    layer = circuits.heron_layered(n, 1, pairs, seed=n)
    instr = []
    for k in (5, 6, min(8, n), 7 if n >= 7 else 5):
        qs = [int(q) for q in rng.choice(n, size=k, replace=False)]
        instr += [["unitary", qs, _haar(rng, k)]] + layer
    instr += [["unitary", [int(q) for q in rng.choice(n, size=5, replace=False)], _haar(rng, 5)]]      # last instruction
    ref = oracle.Program(n, instr, noisy=False).pure_state()
    for reorder in (1, 0):
        E = EmuProgram(PackedCircuit(n, instr, noisy=False), 2, tile_bits=tile_bits, reg_bits=reg_bits, reorder=reorder)
        assert np.abs(E.run_deterministic() - ref).max() < 1e-12
    # noisy trajectories: replay with explicit uniforms (the engine's order; the uniforms stay with their instructions)
    E = EmuProgram(PackedCircuit(n, instr, sq, tq), 0, tile_bits=tile_bits, reg_bits=reg_bits)
    order = E.order()
    U = np.random.default_rng(1).random((3, len(instr)))
    U[1:] = 1 - 0.03 * np.random.default_rng(2).random((2, len(instr))) ** 2
    states, stats = E.mcwf_states(U)
    O = oracle.Program(n, [instr[g] for g in order], sq, tq)
    for t in range(3):
        assert np.abs(states[t] - O.trajectory(U[t][order], mask_fix=True)).max() < TOL
    assert stats["events"] > 0


def test_density_matrix_with_a_five_qubit_unitary(heron_noise):
    sq, tq, _, _ = heron_noise
    n = 5
    rng = np.random.default_rng(3)
    instr = circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n), seed=1) + [["unitary", [3, 0, 4, 1, 2], _haar(rng, 5)]] + \
        circuits.heron_layered(n, 1, circuits.coupled_pairs(tq, n), seed=2)
    E = EmuProgram(PackedCircuit(n, instr, sq, tq), 1, tile_bits=9, reg_bits=4)
    rho = E.run_deterministic().reshape(1 << n, 1 << n).T
    assert np.abs(rho - oracle.density_matrix(n, instr, sq, tq)).max() < TOL
