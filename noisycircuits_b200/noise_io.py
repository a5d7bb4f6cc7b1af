"""
Flat (npz) storage for the noise tables ``QuantumCircuit.__init__`` hands to a solver back-end
(/root/reference/src/NoisyCircuits/QuantumCircuit.py:158-169):

    single_qubit_noise : {qubit: {gate: [ndarray(2,2) complex128, ...]}}         ascending qubit order
    two_qubit_noise    : {gate: {(q0, q1): [ndarray(4,4) complex128, ...]}}      already [0,2,1,3]-permuted
    measurement_error  : {qubit: ndarray(2,2) float64}

The reference rebuilds these with joblib on every construction; here they can be cached on disk and shipped as
fixtures (tests/golden/*.npz), since the reference package itself is not available on the GPU box.
"""
from __future__ import annotations

import numpy as np


def save_noise_tables(path, single_qubit_noise, two_qubit_noise, measurement_error, connectivity=None, coupling_map=None):
    gates = sorted({g for d in single_qubit_noise.values() for g in d} | set(two_qubit_noise))
    gid = {g: i for i, g in enumerate(gates)}
    sq_index, sq_data = [], []
    for q, d in single_qubit_noise.items():
        for g, ops in d.items():
            sq_index.append((int(q), gid[g], len(ops)))
            sq_data.extend(np.asarray(k, np.complex128).reshape(2, 2) for k in ops)
    tq_index, tq_data = [], []
    for g, d in two_qubit_noise.items():
        for (q0, q1), ops in d.items():
            tq_index.append((gid[g], int(q0), int(q1), len(ops)))
            tq_data.extend(np.asarray(k, np.complex128).reshape(4, 4) for k in ops)
    meas_q = np.asarray(list(measurement_error.keys()), np.int32)
    meas = np.asarray([np.asarray(m, np.float64) for m in measurement_error.values()], np.float64).reshape(-1, 2, 2)
    extra = {}
    if coupling_map is not None:
        extra["coupling_map"] = np.asarray(coupling_map, np.int32).reshape(-1, 2)
    np.savez_compressed(
        path, gates=np.asarray(gates), sq_index=np.asarray(sq_index, np.int32).reshape(-1, 3),
        sq_data=np.asarray(sq_data, np.complex128).reshape(-1, 2, 2),
        tq_index=np.asarray(tq_index, np.int32).reshape(-1, 4),
        tq_data=np.asarray(tq_data, np.complex128).reshape(-1, 4, 4), meas_q=meas_q, meas=meas, **extra)


def load_noise_tables(path):
    """-> (single_qubit_noise, two_qubit_noise, measurement_error, coupling_map)"""
    z = np.load(path, allow_pickle=False)
    gates = [str(g) for g in z["gates"]]
    single, two = {}, {}
    pos = 0
    for q, g, cnt in z["sq_index"]:
        single.setdefault(int(q), {})[gates[g]] = [np.array(k) for k in z["sq_data"][pos:pos + cnt]]
        pos += cnt
    pos = 0
    for g, q0, q1, cnt in z["tq_index"]:
        two.setdefault(gates[g], {})[(int(q0), int(q1))] = [np.array(k) for k in z["tq_data"][pos:pos + cnt]]
        pos += cnt
    meas = {int(q): np.array(m) for q, m in zip(z["meas_q"], z["meas"])}
    cmap = [tuple(int(v) for v in p) for p in z["coupling_map"]] if "coupling_map" in z.files else None
    return single, two, meas, cmap
