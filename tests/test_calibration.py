"""Noise tables: the committed fixtures are CPTP; the qiskit-aer-free calibration restatement
(noisycircuits_b200/calibration.py) emits the schema of the shipped Eagle pickle and the reference's BuildModel
accepts it (only where /root/reference exists)."""
import os
import pickle
import sys

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cptp(ops):
    d = ops[0].shape[0]
    s = sum(k.conj().T @ k for k in ops)
    return np.abs(s - np.eye(d)).max()


def test_fixture_channels_are_cptp(heron_noise, eagle_noise):
    for sq, tq, meas, _ in (heron_noise, eagle_noise):
        for q, gates in sq.items():
            for g, ops in gates.items():
                assert _cptp(ops) < 1e-9, (q, g)
        for g, pairs in tq.items():
            for pr, ops in pairs.items():
                assert _cptp(ops) < 1e-9, (g, pr)
        for q, m in meas.items():
            assert np.allclose(m.sum(axis=0), 1) or np.allclose(m.sum(axis=1), 1)


def test_thermal_relaxation_restatement_is_a_channel():
    from noisycircuits_b200.calibration import _Thermal
    for t1, t2, t in ((200e-6, 50e-6, 36e-9), (100e-6, 150e-6, 68e-9), (50e-6, 100e-6, 500e-9)):
        th = _Thermal(t1, t2, t)
        assert abs(sum(p for p, _ in th.terms) - 1) < 1e-12
        if t2 > t1:
            ins = th.terms[0][1][0]
            assert ins["name"] == "kraus"
            ops = [np.array([[complex(*e) for e in row] for row in k]) for k in ins["params"]]
            assert _cptp(ops) < 1e-12
            # process fidelity of the Kraus form = tr(S)/4 with S = sum K* (x) K
            fpro = sum(abs(np.trace(k)) ** 2 for k in ops) / 4
            assert abs(fpro - th.f_pro) < 1e-12
        else:
            assert [i[0]["name"] for _, i in th.terms] == ["id", "z", "reset"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_calibration_schema_matches_eagle_pickle_and_buildmodel_accepts_it():
    from noisycircuits_b200.calibration import create_noise_model
    nm = create_noise_model(os.path.join(REF, "noise_models", "Sample_Noise_Model_Heron_QPU.csv"),
                            [["rz", "rx", "sx", "x"], ["cz", "rzz"]])
    eagle = pickle.load(open(os.path.join(REF, "noise_models", "Noise_Model_Eagle_QPU.pkl"), "rb"))
    keys = lambda errs, typ: {tuple(sorted(e.keys())) for e in errs if e["type"] == typ}
    assert keys(nm["errors"], "qerror") <= keys(eagle["errors"], "qerror")
    assert keys(nm["errors"], "roerror") == keys(eagle["errors"], "roerror")
    names = {i["name"] for e in nm["errors"] if e["type"] == "qerror" for ins in e["instructions"] for i in ins}
    assert names <= {"id", "x", "y", "z", "reset", "kraus", "pauli"}
    for e in nm["errors"]:
        if e["type"] == "qerror":
            assert abs(sum(e["probabilities"]) - 1) < 1e-12
    # the committed Heron tables are what BuildModel makes of this dictionary
    sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(REF, "src")]
    try:
        from NoisyCircuits.utils.BuildQubitGateModel import BuildModel
    except Exception as exc:      # reference extensions not built here
        pytest.skip(f"reference package not importable: {exc}")
    m = BuildModel(noise_model=nm, num_qubits=3, num_cores=1, threshold=1e-6, basis_gates=[["rz", "rx", "sx", "x"], ["cz", "rzz"]], verbose=False)
    single, multi, _meas, _conn = m.build_qubit_gate_model()
    from noisycircuits_b200 import noise_io
    sq, _tq, _m, _c = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
    for q in range(3):
        for g in ("rz", "rx", "sx", "x"):
            a, b = single[q][g]["qubit_channel"], sq[q][g]
            assert len(a) == len(b) and all(np.abs(x - y).max() < 1e-12 for x, y in zip(a, b))


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_install_registers_the_backend_with_the_reference_package():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(REF, "src")]
    try:
        import NoisyCircuits  # noqa: F401
    except Exception as exc:
        pytest.skip(f"reference package not importable: {exc}")
    import noisycircuits_b200
    from NoisyCircuits.QuantumCircuit import QuantumCircuit
    from NoisyCircuits.utils.solvers import load_solver
    assert noisycircuits_b200.install()
    assert "b200" in QuantumCircuit.available_sim_backends
    mod = load_solver("b200")
    for cls in ("RemoteExecutor", "DensityMatrixSolver", "PureStateSolver"):
        assert hasattr(mod, cls)
