"""
Generates the committed fixtures under tests/golden/ from the REFERENCE itself.  Run in the build container
(where /root/reference exists), after `make -C oracle all`:

    python tests/golden/make_fixtures.py

Outputs
  heron20_noise.npz   BuildModel(threshold=1e-6) tables for the first 20 qubits of
                      noise_models/Sample_Noise_Model_Heron_QPU.csv.  The CSV -> noise-dict step is
                      noisycircuits_b200.calibration (restated; qiskit-aer is not installed), the
                      dict -> Kraus-table step is the reference's own BuildModel + convert_matrix_to_little_endian.
  eagle16_noise.npz   same for the first 16 qubits of noise_models/Noise_Model_Eagle_QPU.pkl (pure reference data).
  ref_vectors.npz     known-answer vectors produced by the reference's compiled extensions (oracle/_ref):
                      simulator.simulate_circuit (pure + noisy), simulator_mpi.run_trajectory,
                      measurement_error_applicator.apply_measurement_error; verbatim build on adjacent-pair
                      circuits, patched build (quad-mask fix) on circuits with non-adjacent pairs.
"""
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
_PATHS = [ROOT, os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(REF, "src")]
sys.path[:0] = _PATHS
# joblib/loky workers re-import NoisyCircuits: they need the same search path
os.environ["PYTHONPATH"] = os.pathsep.join(_PATHS + [os.environ.get("PYTHONPATH", "")])

from noisycircuits_b200 import circuits, noise_io                    # noqa: E402
from noisycircuits_b200.calibration import create_noise_model       # noqa: E402


def import_ref(patched):
    """The reference's extensions from oracle/_ref (or oracle/_ref/patched).  CPython caches extension modules
    by name, so the patched set can only be imported in a process that never imported the verbatim set: the
    patched section of main() runs in a subprocess (see `--patched`)."""
    import importlib
    d = os.path.join(ROOT, "oracle", "_ref", "patched" if patched else "")
    if patched:
        sys.path[:] = [p for p in sys.path if p != os.path.join(ROOT, "oracle", "_ref")]
    sys.path.insert(0, d)
    return [importlib.import_module(m) for m in ("simulator", "simulator_mpi", "measurement_error_applicator")]


def build_tables(noise_dict, n, basis_gates):
    from NoisyCircuits.utils.BuildQubitGateModel import BuildModel
    from NoisyCircuits.utils.HelperFunctions import convert_matrix_to_little_endian
    m = BuildModel(noise_model=noise_dict, num_qubits=n, num_cores=4, threshold=1e-6, basis_gates=basis_gates, verbose=False)
    single, multi, meas, _conn = m.build_qubit_gate_model()
    # QuantumCircuit.py:158-169
    sq = {q: {g: p["qubit_channel"] for g, p in gates.items()} for q, gates in single.items()}
    tq = {g: {pair: convert_matrix_to_little_endian(p["qubit_channel"]) for pair, p in pairs.items()} for g, pairs in multi.items()}
    return sq, tq, meas, m.qubit_coupling_map


def jsonable(instr):
    out = []
    for name, qubits, param in instr:
        if name == "unitary":
            U = np.asarray(param, complex)
            out.append([name, list(map(int, qubits)), {"re": U.real.tolist(), "im": U.imag.tolist()}])
        else:
            out.append([name, list(map(int, qubits)), float(param)])
    return out


def random_unitary(rng, d):
    a = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def patched_main(path):
    """Section (d): non-adjacent pairs on the PATCHED build: pure + noisy (noise tables re-keyed onto the
    non-adjacent pairs)."""
    h_sq, h_tq, _h_meas, _ = noise_io.load_noise_tables(os.path.join(HERE, "heron20_noise.npz"))
    psim, psim_mpi, _ = import_ref(patched=True)
    out = {}
    n = 5
    remap = {(0, 2): (0, 1), (4, 1): (2, 1), (3, 0): (3, 2)}
    tq_far = {g: {far: h_tq[g][near] for far, near in remap.items()} for g in ("cz", "rzz")}
    instr = []
    for layer in range(3):
        for q in range(n):
            instr.append([("rx", "sx", "rz", "x")[(q + layer) % 4], [q, q], 0.4 * (q + 1) + layer])
        for k, pr in enumerate(remap):
            instr.append(["cz", list(pr), 0] if (k + layer) % 2 else ["rzz", list(pr), 0.8 - 0.3 * k + layer])
    buf = np.zeros(1 << n, np.complex128)
    psim.simulate_circuit(instr, buf, {}, {}, n, False, 1, True, 1)
    out["pure_far_state"] = buf.copy()
    buf = np.zeros(1 << n, np.complex128)
    psim.simulate_circuit(instr, buf, h_sq, tq_far, n, True, 64, False, 2)
    out["mcwf_far_probs"] = buf.real.copy()
    trajs = []
    for t in range(4):
        buf = np.zeros(1 << n, np.complex128)
        psim_mpi.run_trajectory(instr, buf, h_sq, tq_far, n, 42 + t, 1)
        trajs.append(buf.real.copy())
    out["mcwf_far_traj"] = np.stack(trajs)
    meta = {"n": n, "instr": jsonable(instr), "T": 64, "noise": "heron20",
            "remap": [[list(k), list(v)] for k, v in remap.items()], "traj_seeds": [42 + t for t in range(4)]}
    np.savez(path, meta=json.dumps(meta), **out)


def main():
    heron_gates = [["rz", "rx", "sx", "x"], ["cz", "rzz"]]
    eagle_gates = [["rz", "sx", "x"], ["ecr"]]
    nm = create_noise_model(os.path.join(REF, "noise_models", "Sample_Noise_Model_Heron_QPU.csv"), heron_gates)
    h_sq, h_tq, h_meas, h_map = build_tables(nm, 20, heron_gates)
    noise_io.save_noise_tables(os.path.join(HERE, "heron20_noise.npz"), h_sq, h_tq, h_meas, coupling_map=h_map)
    with open(os.path.join(REF, "noise_models", "Noise_Model_Eagle_QPU.pkl"), "rb") as fh:
        eagle = pickle.load(fh)
    e_sq, e_tq, e_meas, e_map = build_tables(eagle, 16, eagle_gates)
    noise_io.save_noise_tables(os.path.join(HERE, "eagle16_noise.npz"), e_sq, e_tq, e_meas, coupling_map=e_map)

    sim, sim_mpi, mea = import_ref(patched=False)
    rng = np.random.default_rng(2024)
    out, meta = {}, {}

    def run_noisy(s, n, instr, sq, tq, T):
        buf = np.zeros(1 << n, np.complex128)
        s.simulate_circuit(instr, buf, sq, tq, n, True, T, False, 2)
        return buf.real.copy()

    def run_traj(s, n, instr, sq, tq, seed):
        buf = np.zeros(1 << n, np.complex128)
        s.run_trajectory(instr, buf, sq, tq, n, seed, 1)
        return buf.real.copy()

    # (a) every gate kernel, adjacent pairs, verbatim build, noise-free
    n = 5
    instr = []
    for q in range(n):
        instr += [["h", [q, q], 0], ["rx", [q, q], 0.3 + q], ["ry", [q, q], -0.7 * (q + 1)], ["rz", [q, q], 1.1 * q - 2],
                  ["p", [q, q], 0.37 * (q + 1)], ["sx", [q, q], 0], ["x", [q, q], 0]]
    for a, b in [(0, 1), (2, 1), (2, 3), (4, 3)]:
        instr += [["cz", [a, b], 0], ["rzz", [a, b], 0.9 + a], ["ecr", [a, b], 0], ["cx", [b, a], 0], ["swap", [a, b], 0],
                  ["rx", [a, a], 0.2 * b], ["sx", [b, b], 0]]
    instr += [["unitary", [1, 2], random_unitary(rng, 4)], ["unitary", [4, 2, 0], random_unitary(rng, 8)],
              ["unitary", [3], random_unitary(rng, 2)]]
    buf = np.zeros(1 << n, np.complex128)
    sim.simulate_circuit(instr, buf, {}, {}, n, False, 1, True, 1)
    out["pure_adjacent_state"] = buf.copy()
    meta["pure_adjacent"] = {"n": n, "instr": jsonable(instr)}

    # (b) noisy Heron ansatz on the 0-1-2-3 chain, verbatim build
    n = 4
    pairs = circuits.coupled_pairs(h_tq, n)
    instr = circuits.heron_ansatz(n, 3, pairs, seed=7)
    out["mcwf_heron4_probs"] = run_noisy(sim, n, instr, h_sq, h_tq, 64)
    out["mcwf_heron4_traj"] = np.stack([run_traj(sim_mpi, n, instr, h_sq, h_tq, 42 + t) for t in range(6)])
    meta["mcwf_heron4"] = {"n": n, "instr": jsonable(instr), "T": 64, "noise": "heron20", "traj_seeds": [42 + t for t in range(6)]}

    # (c) noisy Eagle layered circuit on the 0-1-2-3 chain, verbatim build
    pairs = circuits.coupled_pairs(e_tq, n, adjacent_only=True)
    instr = circuits.eagle_layered(n, 3, pairs, seed=11)
    out["mcwf_eagle4_probs"] = run_noisy(sim, n, instr, e_sq, e_tq, 64)
    out["mcwf_eagle4_traj"] = np.stack([run_traj(sim_mpi, n, instr, e_sq, e_tq, 42 + t) for t in range(6)])
    meta["mcwf_eagle4"] = {"n": n, "instr": jsonable(instr), "T": 64, "noise": "eagle16", "traj_seeds": [42 + t for t in range(6)]}

    # (d) non-adjacent pairs, PATCHED build, in a subprocess (see import_ref)
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        tmp = os.path.join(td, "far.npz")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--patched", tmp], check=True)
        zf = np.load(tmp)
        for k in zf.files:
            if k != "meta":
                out[k] = zf[k]
        meta["far"] = json.loads(str(zf["meta"]))

    # (e) measurement error
    n = 5
    p = rng.random(1 << n); p /= p.sum()
    q = p.copy()
    mea.apply_measurement_error(q, {k: h_meas[k] for k in range(n)}, list(range(n)), n, 1)
    out["meas_in"], out["meas_out"] = p, q
    meta["meas"] = {"n": n, "noise": "heron20"}

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), meta=json.dumps(meta), **out)
    for k, v in out.items():
        print(k, v.shape, float(np.abs(v).sum()))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--patched":
        patched_main(sys.argv[2])
    else:
        main()
