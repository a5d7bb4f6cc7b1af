"""
Synthetic basis-gate circuits of the shapes BASELINE.json / SURVEY.md §8(d) name, emitted directly in the
instruction-list format the reference's decomposers produce
(/root/reference/src/NoisyCircuits/utils/HeronDecomposition.py:45-146, EagleDecomposition.py:44-118):

    [name, [q, q], theta]        single-qubit gate (the qubit is carried twice)
    [name, [q0, q1], theta]      two-qubit gate, qubits in the noise model's direction

All generators are seeded with ``numpy.random.default_rng(seed)`` and angles are U(-2*pi, 2*pi) as in
docs/source/theory/comparison.rst:89-99.
"""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi


def coupled_pairs(two_qubit_noise: dict, num_qubits: int, gate: str | None = None, adjacent_only: bool = False):
    """Undirected coupled pairs among the first ``num_qubits`` qubits, each in the direction it first appears
    in the noise model (the direction the reference's decomposers emit, HeronDecomposition.py:127-131)."""
    gate = gate or next(iter(two_qubit_noise))
    seen, out = set(), []
    for (a, b) in two_qubit_noise[gate].keys():
        if a >= num_qubits or b >= num_qubits:
            continue
        key = (min(a, b), max(a, b))
        if key in seen:
            continue
        if adjacent_only and abs(a - b) != 1:
            continue
        seen.add(key)
        out.append((int(a), int(b)))
    return out


def heron_layered(num_qubits: int, depth: int, pairs, seed: int = 42):
    """Random layered Heron circuit (configs C2/C5): per layer one of {x, sx, rx(t), rz(t)} on every qubit, then
    cz or rzz(t) on every coupled pair."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(depth):
        for q in range(num_qubits):
            g = ("x", "sx", "rx", "rz")[rng.integers(4)]
            th = float(rng.uniform(-TWO_PI, TWO_PI)) if g in ("rx", "rz") else 0
            out.append([g, [q, q], th])
        for (a, b) in pairs:
            if rng.integers(2):
                out.append(["cz", [a, b], 0])
            else:
                out.append(["rzz", [a, b], float(rng.uniform(-TWO_PI, TWO_PI))])
    return out


def eagle_layered(num_qubits: int, depth: int, pairs, seed: int = 42):
    """Random layered Eagle circuit (config C4): per layer one of {x, sx, rz(t)} on every qubit, then ecr on every
    coupled pair."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(depth):
        for q in range(num_qubits):
            g = ("x", "sx", "rz")[rng.integers(3)]
            th = float(rng.uniform(-TWO_PI, TWO_PI)) if g == "rz" else 0
            out.append([g, [q, q], th])
        for (a, b) in pairs:
            out.append(["ecr", [a, b], 0])
    return out


def heron_ansatz(num_qubits: int, layers: int, pairs, seed: int = 42):
    """Config C1: L layers of {rx, rz, sx on every qubit; cz or rzz(t) on every coupled pair}."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(layers):
        for q in range(num_qubits):
            out.append(["rx", [q, q], float(rng.uniform(-TWO_PI, TWO_PI))])
            out.append(["rz", [q, q], float(rng.uniform(-TWO_PI, TWO_PI))])
            out.append(["sx", [q, q], 0])
        for (a, b) in pairs:
            if rng.integers(2):
                out.append(["cz", [a, b], 0])
            else:
                out.append(["rzz", [a, b], float(rng.uniform(-TWO_PI, TWO_PI))])
    return out


def qnn_circuit(num_qubits: int, layers: int, pairs, seed: int = 42):
    """Config C3 (examples/quantum_neural_networks.ipynb cells 17-18, 23): per layer RY(w) on every qubit -- which
    the Heron decomposer emits as x, sx, rz(-w), sx (HeronDecomposition.py:75-82) --, a CZ chain, RX(x) on every
    qubit."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(layers):
        for q in range(num_qubits):
            w = float(rng.uniform(-TWO_PI, TWO_PI))
            out += [["x", [q, q], 0], ["sx", [q, q], 0], ["rz", [q, q], -w], ["sx", [q, q], 0]]
        for (a, b) in pairs:
            out.append(["cz", [a, b], 0])
        for q in range(num_qubits):
            out.append(["rx", [q, q], float(rng.uniform(-TWO_PI, TWO_PI))])
    return out
