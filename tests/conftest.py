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
