"""The CPU oracle (oracle/) against the known-answer vectors generated from the reference's own compiled
extensions (tests/golden/ref_vectors.npz, made by tests/golden/make_fixtures.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import oracle
from conftest import far_tables

TOL = 1e-12


def test_pure_state_all_gate_kernels(golden):
    z, meta = golden
    m = meta["pure_adjacent"]
    P = oracle.Program(m["n"], m["instr"], noisy=False)
    for fix in (False, True):            # adjacent pairs: both quad maps agree
        assert np.abs(P.pure_state(mask_fix=fix) - z["pure_adjacent_state"]).max() < TOL


def test_pure_state_non_adjacent_pairs_needs_mask_fix(golden):
    z, meta = golden
    m = meta["far"]
    P = oracle.Program(m["n"], m["instr"], noisy=False)
    assert np.abs(P.pure_state(mask_fix=True) - z["pure_far_state"]).max() < TOL
    # the verbatim quad map is a different (wrong) function on |q1-q2| >= 2 -- SURVEY.md section 0.3
    assert np.abs(P.pure_state(mask_fix=False) - z["pure_far_state"]).max() > 1e-3


@pytest.mark.parametrize("case,noise", [("mcwf_heron4", "heron"), ("mcwf_eagle4", "eagle")])
def test_mcwf_reference_seeding(golden, heron_noise, eagle_noise, case, noise):
    z, meta = golden
    m = meta[case]
    sq, tq, _meas, _ = heron_noise if noise == "heron" else eagle_noise
    P = oracle.Program(m["n"], m["instr"], sq, tq)
    # single trajectories, seeds 42.. (SimulatorMPI.cpp:40-59)
    for row, seed in zip(z[case + "_traj"], m["traj_seeds"]):
        st = P.trajectory(P.reference_uniforms(seed), mask_fix=False)
        assert np.abs(np.abs(st) ** 2 - row).max() < TOL
    # trajectory average with the reference's seed = 42 + t (Simulator.cpp:118-133)
    probs = P.mcwf(m["T"], seed0=42, mask_fix=False, threads=2)
    assert np.abs(probs - z[case + "_probs"]).max() < TOL
    assert abs(probs.sum() - 1) < 1e-12


def test_mcwf_non_adjacent_pairs(golden, heron_noise):
    z, meta = golden
    m = meta["far"]
    P = oracle.Program(m["n"], m["instr"], heron_noise[0], far_tables(heron_noise, meta))
    for row, seed in zip(z["mcwf_far_traj"], m["traj_seeds"]):
        st = P.trajectory(P.reference_uniforms(seed), mask_fix=True)
        assert np.abs(np.abs(st) ** 2 - row).max() < TOL
    assert np.abs(P.mcwf(m["T"], seed0=42, mask_fix=True) - z["mcwf_far_probs"]).max() < TOL


def test_measurement_error(golden, heron_noise):
    z, meta = golden
    n = meta["meas"]["n"]
    meas = {q: heron_noise[2][q] for q in range(n)}
    p = z["meas_in"].copy()
    oracle.measurement_error(p, meas, list(range(n)), n)
    assert np.abs(p - z["meas_out"]).max() < 1e-15


def test_explicit_uniforms_equal_seeded_run(golden, heron_noise):
    """Replay mode (explicit u per instruction) is the same function as the seeded mode."""
    z, meta = golden
    m = meta["mcwf_heron4"]
    P = oracle.Program(m["n"], m["instr"], heron_noise[0], heron_noise[1])
    U = np.stack([P.reference_uniforms(42 + t) for t in range(m["T"])])
    assert np.abs(P.mcwf(m["T"], uniforms=U, mask_fix=False) - z["mcwf_heron4_probs"]).max() < TOL


def test_density_matrix_restatement_matches_mcwf_statistically(golden, heron_noise):
    """the DM semantics are pinned to the MCWF path statistically (TV <= 3/sqrt(N))."""
    _z, meta = golden
    m = meta["mcwf_heron4"]
    sq, tq, _, _ = heron_noise
    dm = oracle.density_matrix_probs(m["n"], m["instr"], sq, tq)
    assert abs(dm.sum() - 1) < 1e-12
    N = 4000
    P = oracle.Program(m["n"], m["instr"], sq, tq)
    mc = P.mcwf(N, seed0=42, mask_fix=False, threads=4)
    tv = 0.5 * np.abs(mc - dm).sum()
    assert tv <= 3 / np.sqrt(N)
