import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def instr_from_json(items):
    out = []
    for name, qubits, param in items:
        if name == "unitary":
            param = np.asarray(param["re"]) + 1j * np.asarray(param["im"])
        out.append([name, list(qubits), param])
    return out


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLDEN, "ref_vectors.npz"))
    meta = json.loads(str(z["meta"]))
    for m in meta.values():
        if "instr" in m:
            m["instr"] = instr_from_json(m["instr"])
    return z, meta


@pytest.fixture(scope="session")
def heron_noise():
    from noisycircuits_b200 import noise_io
    return noise_io.load_noise_tables(os.path.join(GOLDEN, "heron20_noise.npz"))


@pytest.fixture(scope="session")
def eagle_noise():
    from noisycircuits_b200 import noise_io
    return noise_io.load_noise_tables(os.path.join(GOLDEN, "eagle16_noise.npz"))


def far_tables(heron_noise, meta):
    """Noise tables of the `far` golden case: Heron tables re-keyed onto non-adjacent pairs."""
    _sq, tq, _m, _c = heron_noise
    remap = {tuple(k): tuple(v) for k, v in meta["far"]["remap"]}
    return {g: {far: tq[g][near] for far, near in remap.items()} for g in ("cz", "rzz")}


def random_channel(rng, d, kind):
    """A CPTP Kraus list of dimension d: Pauli mixture, amplitude damping (non-unitary base branch, state-dependent
    branch probabilities) or the blocks of a random isometry (dense operators, nothing diagonal)."""
    I2, X, Y, Z = np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
    if kind == "pauli":
        paulis = [I2, X, Y, Z] if d == 2 else [np.kron(a, b) for a in (I2, X, Y, Z) for b in (I2, X, Y, Z)]
        pick = [0] + list(rng.choice(np.arange(1, len(paulis)), size=min(3, len(paulis) - 1), replace=False))
        w = np.array([0.9] + list(rng.random(len(pick) - 1) + 0.1))
        w[1:] *= 0.1 / w[1:].sum()
        return [np.sqrt(wk) * paulis[k].astype(complex) for wk, k in zip(w, pick)]
    if kind == "damping":
        g = 0.02 + 0.1 * rng.random()
        k0, k1 = np.diag([1.0, np.sqrt(1 - g)]).astype(complex), np.array([[0, np.sqrt(g)], [0, 0]], complex)
        return [k0, k1] if d == 2 else [np.kron(a, b) for a in (k0, k1) for b in (k0, k1)]
    m = 3
    q, _ = np.linalg.qr(rng.normal(size=(d * m, d)) + 1j * rng.normal(size=(d * m, d)))
    return [np.ascontiguousarray(q[i * d:(i + 1) * d]) for i in range(m)]


def random_noisy_circuit(rng, n, length=None, lean=False):
    """(instruction list, single-qubit noise, two-qubit noise): every gate name of the reference
    (FunctionMapper.hpp:25-79) with random angles, two-qubit gates on random (also non-adjacent) ordered pairs, noise-free
    1-2 qubit unitaries, a random channel kind per (gate, qubits)."""
    # lean: only what the lean kernels run (no swap, no dense Kraus operators): the direct kernel, fused rotations and
    # in-line jumps are exercised instead of the dense 4x4 kernels
    kinds = ["pauli", "damping"] if lean else ["pauli", "damping", "generic"]
    one_q = ["x", "sx", "rz", "rx", "ry", "h", "p"]
    two_q = ["cz", "ecr", "rzz", "cx"] if lean else ["cz", "ecr", "rzz", "cx", "swap"]
    sq = {q: {g: random_channel(rng, 2, kinds[int(rng.integers(0, len(kinds)))]) for g in one_q} for q in range(n)}
    tq = {g: {} for g in two_q}
    instr = []
    for _ in range(int(length or rng.integers(15, 40))):
        r = rng.random()
        if r < 0.55 or n < 2:
            q = int(rng.integers(0, n))
            instr.append([one_q[int(rng.integers(0, len(one_q)))], [q, q], float(rng.uniform(-6, 6))])
        elif r < 0.92:
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            g = two_q[int(rng.integers(0, len(two_q)))]
            if (a, b) not in tq[g]:
                tq[g][(a, b)] = random_channel(rng, 4, kinds[int(rng.integers(0, len(kinds)))])
            instr.append([g, [a, b], float(rng.uniform(-6, 6))])
        else:
            k = 1 if lean else int(rng.integers(1, 3))
            qs = [int(x) for x in rng.choice(n, size=k, replace=False)]
            u = np.linalg.qr(rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k)))[0]
            instr.append(["unitary", qs, u])
    return instr, sq, tq
