"""The reference's own `QuantumCircuit` class with sim_backend="b200" (QuantumCircuit.py:100-223, 353-446).

/root/reference is not on the GPU box, so the class-level check is split in two halves that meet at committed golden
vectors (tests/golden/class_vectors.npz, made by tests/golden/make_class_fixture.py from the reference class with its
own `custom` back-end):
  * here (CPU, reference checkout present): install() registers the back-end, the class builds the circuit through its
    Heron decomposition front-end, and execute() hands EXACTLY the golden instruction list, the constructor's noise tables
    and the caller's arguments to noisycircuits_b200.backend.execute;
  * on the GPU: backend.execute with those inputs and the reference's mt19937_64 streams returns the reference's
    execute() output (1e-10), for all qubits and for a measured subset; run_pure_state's output likewise (1e-12).
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _golden():
    from noisycircuits_b200 import noise_io
    z = np.load(os.path.join(GOLDEN, "class_vectors.npz"))
    meta = json.loads(str(z["meta"]))
    instr = [[g, list(q), 0 if p is None else p] for g, q, p in meta["instr"]]
    sq, tq, meas, _ = noise_io.load_noise_tables(os.path.join(GOLDEN, "class_noise.npz"))
    return z, meta, instr, sq, tq, meas


@pytest.mark.gpu
def test_b200_backend_returns_the_reference_class_outputs():
    from noisycircuits_b200 import backend
    from noisycircuits_b200.solvers import PureStateSolver
    z, meta, instr, sq, tq, meas = _golden()
    n, T = meta["n"], meta["T"]
    out = backend.execute(n, instr, sq, tq, meas, list(range(n)), T, rng="reference")
    assert np.abs(out - z["execute_all"]).max() < 1e-10
    sub = backend.execute(n, instr, sq, tq, meas, [0, 1, 2], T, rng="reference")
    assert np.abs(sub - z["execute_012"]).max() < 1e-10
    p = PureStateSolver(n, instr, 1, False).solve()
    big_endian = p.reshape([2] * n).transpose(list(range(n))[::-1]).reshape(-1)          # QuantumCircuit.py:574
    assert np.abs(big_endian - z["pure_probs"]).max() < 1e-12


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_reference_class_hands_the_golden_inputs_to_the_b200_backend(monkeypatch):
    paths = [os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(REF, "src")]
    sys.path[:0] = paths
    monkeypatch.setenv("PYTHONPATH", os.pathsep.join(paths + [ROOT, os.environ.get("PYTHONPATH", "")]))    # joblib workers
    try:
        from NoisyCircuits.QuantumCircuit import QuantumCircuit
    except Exception as exc:
        pytest.skip(f"reference package not importable: {exc}")
    import noisycircuits_b200
    from noisycircuits_b200 import backend
    from noisycircuits_b200.calibration import create_noise_model
    sys.path.insert(0, GOLDEN)
    import make_class_fixture as fx
    z, meta, instr, sq, tq, meas = _golden()
    assert noisycircuits_b200.install()
    nm = create_noise_model(os.path.join(REF, "noise_models", "Sample_Noise_Model_Heron_QPU.csv"), fx.HERON_GATES)
    qc = QuantumCircuit(num_qubits=meta["n"], noise_model=nm, backend_qpu_type="heron", basis_gates=fx.HERON_GATES, use_fractional=True,
                        sim_backend="b200", threshold=meta["threshold"], verbose=False)
    assert qc.solver is sys.modules["noisycircuits_b200"]
    fx.build_circuit(qc)
    got = [[g, list(q), 0 if p is None else float(p)] for g, q, p in qc.instruction_list]
    assert got == instr                                   # the decomposition front-end is the reference's, unchanged
    for q in range(meta["n"]):                            # ... and so are the tables its constructor prepared
        for g, ops in sq[q].items():
            assert all(np.array_equal(a, b) for a, b in zip(qc.single_qubit_error[q][g], ops))
    for g, d in tq.items():
        for pair, ops in d.items():
            assert all(np.array_equal(a, b) for a, b in zip(qc.two_qubit_error[g][pair], ops))
    seen = {}

    def fake_execute(num_qubits, instruction_list, single_qubit_noise, two_qubit_noise, measurement_error, qubits, num_trajectories, num_cores, **kw):
        seen.update(n=num_qubits, instr=instruction_list, sq=single_qubit_noise, tq=two_qubit_noise, meas=measurement_error, qubits=qubits,
                    T=num_trajectories, kw=kw)
        return "sentinel"
    with pytest.raises(TypeError):                        # the reference's argument checks (QuantumCircuit.py:391-409) come first
        qc.execute(qubits="012", num_trajectories=10)
    monkeypatch.setattr(backend, "execute", fake_execute)
    assert qc.execute(qubits=[0, 1, 2], num_trajectories=meta["T"], num_cores=2) == "sentinel"
    assert seen["n"] == meta["n"] and seen["instr"] is qc.instruction_list and seen["sq"] is qc.single_qubit_error
    assert seen["tq"] is qc.two_qubit_error and seen["meas"] is qc.measurement_error and seen["qubits"] == [0, 1, 2] and seen["T"] == meta["T"]
    # a circuit object of another back-end is left alone
    qc2 = QuantumCircuit(num_qubits=2, noise_model=nm, backend_qpu_type="heron", basis_gates=fx.HERON_GATES, sim_backend="custom",
                         threshold=1e-6, verbose=False)
    qc2.H(0)
    qc2.CX(0, 1)
    ref = qc2.execute(qubits=[0, 1], num_trajectories=50, num_cores=1)
    assert ref.shape == (4,) and abs(ref.sum() - 1) < 1e-9 and not seen.get("n") == 2


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_noise_operator_cache_in_front_of_the_reference_builder(tmp_path, monkeypatch):
    """install(cache_noise_tables=True): the second construction of a QuantumCircuit on the same noise model takes its
    noise operators from the content-addressed on-disk cache instead of re-running BuildModel (QuantumCircuit.py:149-157)."""
    paths = [os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(REF, "src")]
    sys.path[:0] = paths
    monkeypatch.setenv("PYTHONPATH", os.pathsep.join(paths + [ROOT, os.environ.get("PYTHONPATH", "")]))
    monkeypatch.setenv("NCB_CACHE_DIR", str(tmp_path))
    try:
        from NoisyCircuits.QuantumCircuit import QuantumCircuit
    except Exception as exc:
        pytest.skip(f"reference package not importable: {exc}")
    import noisycircuits_b200
    from noisycircuits_b200 import cache
    from noisycircuits_b200.calibration import create_noise_model
    gates = [["rz", "rx", "sx", "x"], ["cz", "rzz"]]
    assert noisycircuits_b200.install(cache_noise_tables=True)
    nm = create_noise_model(os.path.join(REF, "noise_models", "Sample_Noise_Model_Heron_QPU.csv"), gates)
    before = dict(cache.stats)
    kw = dict(num_qubits=4, noise_model=nm, backend_qpu_type="heron", basis_gates=gates, sim_backend="b200", threshold=1e-6, verbose=False)
    a = QuantumCircuit(**kw)
    b = QuantumCircuit(**kw)
    assert cache.stats["noise_misses"] == before["noise_misses"] + 1 and cache.stats["noise_hits"] == before["noise_hits"] + 1
    assert len([f for f in os.listdir(tmp_path) if f.startswith("noise-")]) == 1
    for q in a.single_qubit_error:
        for g, ops in a.single_qubit_error[q].items():
            assert all(np.array_equal(x, y) for x, y in zip(ops, b.single_qubit_error[q][g]))
    for g in a.two_qubit_error:
        for pair, ops in a.two_qubit_error[g].items():
            assert all(np.array_equal(x, y) for x, y in zip(ops, b.two_qubit_error[g][pair]))
    assert a.connectivity == b.connectivity
    b.H(0); b.CX(0, 1)                                     # the decomposer got its coupling map from the cached entry
    assert len(b.instruction_list) > 0
    c = QuantumCircuit(**dict(kw, threshold=1e-8))         # another threshold: another entry
    assert cache.stats["noise_misses"] == before["noise_misses"] + 2 and c.threshold == 1e-8
