#!/usr/bin/env python
"""Golden vectors for the CLASS-level drop-in test (needs /root/reference; run here, the outputs are committed).

The reference's own `QuantumCircuit` (sim_backend="custom"; its native extensions built from its sources, oracle/_ref)
builds a circuit through its Heron decomposition front-end and executes it.  Stored in tests/golden/class_vectors.npz +
class_noise.npz: the instruction list the decomposer emitted, the noise tables the constructor prepared
(QuantumCircuit.py:149-169), and what execute() / run_pure_state() returned.  tests/test_reference_class.py replays them:
on the GPU through noisycircuits_b200.backend.execute with rng="reference" (the engine must return the reference's
numbers), here on the CPU by checking that QuantumCircuit(sim_backend="b200") hands exactly these inputs to that function.

usage: python tests/golden/make_class_fixture.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
_PATHS = [ROOT, os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(REF, "src")]
sys.path[:0] = _PATHS
os.environ["PYTHONPATH"] = os.pathsep.join(_PATHS + [os.environ.get("PYTHONPATH", "")])

from noisycircuits_b200 import noise_io                              # noqa: E402
from noisycircuits_b200.calibration import create_noise_model       # noqa: E402

HERON_GATES = [["rz", "rx", "sx", "x"], ["cz", "rzz"]]
N = 5


def build_circuit(qc):
    """User-level gates (Decomposition.py:50-817) on index-adjacent pairs of the line 0-1-2-3-4: every one- and two-qubit
    entry point the Heron decomposer offers."""
    qc.H(0)
    qc.CX(0, 1)
    qc.RY(0.37, 2)
    qc.RX(-1.2, 3)
    qc.RZ(0.81, 4)
    qc.CZ(1, 2)
    qc.RZZ(0.6, 2, 3)
    qc.SX(4)
    qc.X(0)
    qc.Y(1)
    qc.Z(2)
    qc.CY(3, 4)
    qc.SWAP(1, 2)
    qc.CRX(0.4, 0, 1)
    qc.CRZ(-0.9, 2, 3)
    qc.CRY(1.1, 3, 4)
    qc.RXX(0.3, 0, 1)
    qc.RYY(-0.5, 1, 2)
    qc.H(3)
    qc.CX(3, 2)


def jsonable(instr):
    out = []
    for name, qubits, param in instr:
        out.append([str(name), [int(q) for q in qubits], None if param is None else float(param)])
    return out


def main():
    from NoisyCircuits.QuantumCircuit import QuantumCircuit
    nm = create_noise_model(os.path.join(REF, "noise_models", "Sample_Noise_Model_Heron_QPU.csv"), HERON_GATES)
    qc = QuantumCircuit(num_qubits=N, noise_model=nm, backend_qpu_type="heron", basis_gates=HERON_GATES, use_fractional=True,
                        sim_backend="custom", threshold=1e-6, verbose=False)
    build_circuit(qc)
    instr = qc.instruction_list
    out = {}
    T = 300
    out["execute_all"] = qc.execute(qubits=list(range(N)), num_trajectories=T, num_cores=2)
    out["execute_012"] = qc.execute(qubits=[0, 1, 2], num_trajectories=T, num_cores=2)
    out["pure_probs"] = qc.run_pure_state(qubits=list(range(N)))
    meta = {"n": N, "T": T, "instr": jsonable(instr), "threshold": 1e-6, "basis_gates": HERON_GATES,
            "note": "reference QuantumCircuit(sim_backend='custom').execute / run_pure_state, seeds 42 + t"}
    noise_io.save_noise_tables(os.path.join(HERE, "class_noise.npz"), qc.single_qubit_error, qc.two_qubit_error, qc.measurement_error)
    np.savez_compressed(os.path.join(HERE, "class_vectors.npz"), meta=json.dumps(meta), **out)
    print("instructions:", len(instr), " sum(execute_all) =", out["execute_all"].sum())


if __name__ == "__main__":
    main()
