"""
Calibration CSV -> qiskit-aer style noise dictionary, without qiskit-aer.

This restates what the reference's ``CreateNoiseModel.create_noise_model`` computes
(/root/reference/src/NoisyCircuits/utils/CreateNoiseModel.py:125-288) for a well-formed calibration table:
per qubit and single-qubit basis gate a ``depolarizing_error(p, 1)`` composed with a
``thermal_relaxation_error(t1, t2, gate_time)``, per coupled pair and two-qubit basis gate a
``depolarizing_error(p, 2)`` composed with the tensor product of the two qubits' thermal relaxation errors,
and a read-out error per qubit, serialised in the ``NoiseModel.to_dict`` schema that the reference's
``BuildModel`` parses (same schema as noise_models/Noise_Model_Eagle_QPU.pkl).

The arithmetic that lives inside qiskit-aer (an unpinned, un-vendored dependency of the reference, not
installed offline) is restated from its published definitions:
  * thermal relaxation, T2 <= T1: mixture  id / z / reset  with
        p_reset = 1 - exp(-t/T1),  p_z = (1 - p_reset) (1 - exp(-t (1/T2 - 1/T1))) / 2,  p_id = 1 - p_z - p_reset
  * thermal relaxation, T2 > T1: Kraus form of the Choi matrix
        [[1, 0, 0, e2], [0, 0, 0, 0], [0, 0, p_reset, 0], [e2, 0, 0, 1 - p_reset]],  e2 = exp(-t/T2)
  * depolarizing_error(p, n): identity with 1 - p (4^n - 1)/4^n, every other Pauli with p/4^n
  * average gate fidelity F = (d F_pro + 1)/(d + 1), F_pro = <<1|Choi|1>>/d^2 (multiplicative over tensor products)
  * depolarizing probability from the reported gate error (CreateNoiseModel.py:98-123)
  * ``a.compose(b)``: all (a_i, b_j) pairs, a outer; ``a.tensor(b)``: a on the higher qubit, a outer, b's
    instructions listed first; zero-probability terms are dropped.
Parity with qiskit-aer's own output is UNPINNED offline (no qiskit-aer here); the structure is checked against the
shipped Eagle pickle and the resulting channels are checked to be CPTP (tests/test_calibration.py).

Only the regular branch of the reference (every two-qubit gate error column lists the same neighbours as the
"Gate Length (ns)" column, CreateNoiseModel.py:176-188) is implemented; irregular rows raise ValueError.
"""
from __future__ import annotations

import csv
import itertools
import math
import uuid

import numpy as np


def _depolarizing_probability(gate_error: float, f_thermal: float, dim: int) -> float:
    # CreateNoiseModel.py:119-123
    p = dim * (gate_error - (1 - f_thermal)) / (dim * f_thermal - 1)
    n = int(round(math.log2(dim)))
    return min(max(0.0, p), 4 ** n / (4 ** n - 1))


class _Thermal:
    """Single-qubit thermal relaxation error as a list of (probability, [instruction dicts on qubit 0])."""

    def __init__(self, t1: float, t2: float, time: float):
        p_reset = 1.0 - math.exp(-time / t1)
        e2 = math.exp(-time / t2)
        self.f_pro = (2.0 + 2.0 * e2 - p_reset) / 4.0
        if t2 > t1:
            # 2x2 block of the Choi matrix on span{|00>>, |11>>}
            blk = np.array([[1.0, e2], [e2, 1.0 - p_reset]])
            w, v = np.linalg.eigh(blk)
            ops = []
            for val, vec in sorted(zip(w, v.T), key=lambda t: -t[0]):
                if abs(val) > 1e-10:
                    ops.append(math.sqrt(val) * np.diag(vec).astype(complex))
            if p_reset > 1e-10:
                ops.append(math.sqrt(p_reset) * np.array([[0, 1], [0, 0]], complex))
            params = [[[[float(e.real), float(e.imag)] for e in row] for row in k] for k in ops]
            self.terms = [(1.0, [{"name": "kraus", "qubits": [0], "params": params}])]
        else:
            p_z = (1.0 - p_reset) * (1.0 - math.exp(-time * (1.0 / t2 - 1.0 / t1))) / 2.0
            p_id = 1.0 - p_z - p_reset
            terms = [(p_id, "id"), (p_z, "z"), (p_reset, "reset")]
            self.terms = [(p, [{"name": nm, "qubits": [0]}]) for p, nm in terms if p > 0.0]


def _shift(instrs, qubit):
    out = []
    for ins in instrs:
        d = dict(ins)
        d["qubits"] = [qubit]
        out.append(d)
    return out


def create_noise_model(calibration_data_file: str, basis_gates) -> dict:
    """Noise dictionary for ``BuildModel`` from a calibration CSV (see module docstring)."""
    with open(calibration_data_file, newline="") as fh:
        rows = list(csv.DictReader(fh))
    cols = rows[0].keys()
    for col in ["Qubit", "T1 (us)", "T2 (us)", "Prob meas 0 prep 1", "Prob meas 1 prep 0",
                "Single Qubit Gate Length (ns)", "Gate Length (ns)"] + [f"{g.lower()} Error" for gs in basis_gates for g in gs]:
        if col not in cols:
            raise ValueError(f"CSV file is missing required column '{col}'.")
    qubits = [int(r["Qubit"]) for r in rows]
    t1 = [float(r["T1 (us)"]) * 1e-6 for r in rows]
    t2 = [float(r["T2 (us)"]) * 1e-6 for r in rows]
    for i in range(len(rows)):
        if t2[i] > 2 * t1[i]:                                   # CreateNoiseModel.py:137-141
            t2[i] = 2 * t1[i]
    errors = []

    def add(op, gate_qubits, terms):
        terms = [(p, ins) for p, ins in terms if p > 0.0]
        errors.append({"type": "qerror", "id": uuid.uuid4().hex, "operations": [op], "gate_qubits": [gate_qubits],
                       "probabilities": [p for p, _ in terms], "instructions": [ins for _, ins in terms]})

    for i, row in enumerate(rows):
        p01 = float(row["Prob meas 0 prep 1"]); p10 = float(row["Prob meas 1 prep 0"])
        errors.append({"type": "roerror", "operations": ["measure"],
                       "probabilities": [[1 - p10, p10], [p01, 1 - p01]], "gate_qubits": [[qubits[i]]]})
        t_1q = int(row["Single Qubit Gate Length (ns)"]) * 1e-9
        for gate in basis_gates[0]:
            raw = row[f"{gate} Error"]
            gate_error = 1.0 if raw in ("", "nan") else float(raw)
            th = _Thermal(t1[i], t2[i], t_1q)
            f_th = (2 * th.f_pro + 1) / 3
            p = _depolarizing_probability(gate_error, f_th, 2)
            depol = [(1 - p * 3 / 4, "id")] + [(p / 4, nm) for nm in ("x", "y", "z")]
            terms = [(pd * pt, [{"name": nm, "qubits": [0]}] + ti)
                     for pd, nm in depol if pd > 0.0 for pt, ti in th.terms]
            add(gate, [qubits[i]], terms)
        lengths = row["Gate Length (ns)"].split(";")
        for gate in basis_gates[1]:
            raw = row[f"{gate} Error"]
            entries = raw.split(";") if raw not in ("", "nan") else []
            if len(entries) != len(lengths):
                raise ValueError(f"irregular two-qubit calibration row for qubit {qubits[i]} gate {gate}: "
                                 "only the regular branch of CreateNoiseModel is restated")
            for ent, gl in zip(entries, lengths):
                tq, terr = ent.split(":"); _, tlen = gl.split(":")
                tq = int(tq); gate_time = int(tlen) * 1e-9
                th_hi = _Thermal(t1[tq], t2[tq], gate_time)      # .tensor receiver: error qubit 1 -> target
                th_lo = _Thermal(t1[i], t2[i], gate_time)        # error qubit 0 -> row qubit
                f_th = (4 * th_hi.f_pro * th_lo.f_pro + 1) / 5
                p = _depolarizing_probability(float(terr), f_th, 4)
                thermal = [(ph * pl, _shift(il, 0) + _shift(ih, 1)) for ph, ih in th_hi.terms for pl, il in th_lo.terms]
                labels = ["".join(t) for t in itertools.product("IXYZ", repeat=2)]
                depol = [(1 - p * 15 / 16, labels[0])] + [(p / 16, lb) for lb in labels[1:]]
                terms = [(pd * pt, [{"name": "pauli", "qubits": [0, 1], "params": [lb]}] + ti)
                         for pd, lb in depol if pd > 0.0 for pt, ti in thermal]
                add(gate, [qubits[i], qubits[tq]], terms)
    return {"errors": errors}
