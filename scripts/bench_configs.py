#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations on one GPU (parity of each is covered by tests/).
Prints one JSON line per configuration.  Usage: python scripts/bench_configs.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisycircuits_b200 import circuits, noise_io                      # noqa: E402
from noisycircuits_b200.solvers import DensityMatrixSolver, RemoteExecutor  # noqa: E402

quick = "--quick" in sys.argv
G = os.path.join(ROOT, "tests", "golden")
h_sq, h_tq, _hm, _ = noise_io.load_noise_tables(os.path.join(G, "heron20_noise.npz"))
e_sq, e_tq, _em, _ = noise_io.load_noise_tables(os.path.join(G, "eagle16_noise.npz"))


def mcwf(name, n, instr, sq, tq, T, reps=2, **opts):
    ex = RemoteExecutor(n, sq, tq, 1, **opts)
    ex.run(min(T, 64), instr)                      # compile + upload + warm-up
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        p = ex.run(T, instr)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ex.last_stats
    print(json.dumps({"config": name, "qubits": n, "instructions": len(instr), "trajectories": T, "seconds": best,
                      "trajectories_per_s": T / best, "kernel_ms": st["kernel_ms"], "sweeps_per_trajectory": st["sweeps"] / T,
                      "events_per_trajectory": st["events"] / T, "hard_events_per_trajectory": st["hard_events"] / T,
                      "norm": float(p.sum())}), flush=True)


def dm(name, n, instr, sq, tq):
    s = DensityMatrixSolver(n, sq, tq, instr, 1)
    s.solve(list(range(n)))
    t0 = time.perf_counter()
    p = s.solve(list(range(n)))
    dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "qubits": n, "instructions": len(instr), "seconds": dt, "state_bytes": 16 * 4 ** n,
                      "algorithmic_GBps": len(instr) * 32 * 4 ** n / dt / 1e9, "trace": float(p.sum())}), flush=True)


# C1: 4-qubit Heron CZ/RZZ ansatz, 1000 trajectories
mcwf("C1 4q Heron ansatz x1000", 4, circuits.heron_ansatz(4, 4, circuits.coupled_pairs(h_tq, 4), seed=42), h_sq, h_tq, 1000)
# C2: 12-qubit density matrix, layered Heron circuit
dm("C2 12q density matrix depth 10", 12, circuits.heron_layered(12, 4 if quick else 10, circuits.coupled_pairs(h_tq, 12), seed=42), h_sq, h_tq)
# C3: 12-qubit QNN-style circuit, 10^4 trajectories
mcwf("C3 12q QNN x1e4", 12, circuits.qnn_circuit(12, 4, circuits.coupled_pairs(h_tq, 12, adjacent_only=True), seed=42), h_sq, h_tq, 2000 if quick else 10000)
# C4: 16-qubit Eagle ECR/SX/RZ depth 40
# (C4 / C5 stand for 10^5-trajectory runs: RemoteExecutor would pick the literal-coefficient generated kernels there by itself)
mcwf("C4 16q Eagle depth 40", 16, circuits.eagle_layered(16, 40, circuits.coupled_pairs(e_tq, 16), seed=42), e_sq, e_tq, 1184 if quick else 4736, specialize=True)
# C5: 20-qubit Heron depth 50
mcwf("C5 20q Heron depth 50", 20, circuits.heron_layered(20, 50, circuits.coupled_pairs(h_tq, 20), seed=42), h_sq, h_tq, 296 if quick else 1184, specialize=True)


# C3 sweep: 20 parameter sets of the QNN skeleton (a training loop's inner evaluation): loop of run() vs run_many()
def sweep(name, n, lists, sq, tq, T):
    ex = RemoteExecutor(n, sq, tq, 1)
    ex.run(64, lists[0])
    t0 = time.perf_counter()
    a = [ex.run(T, instr) for instr in lists]
    t_loop = time.perf_counter() - t0
    ex.run_many(T, lists[:6])                      # (the engine objects of the pipeline created)
    t0 = time.perf_counter()
    b = ex.run_many(T, lists)
    t_many = time.perf_counter() - t0
    # the same sweep as a parameter batch over one gate skeleton (angles as an array, no Python loop per circuit)
    thetas = np.array([[0.0 if p is None else float(p) for _g, _q, p in instr] for instr in lists])
    t0 = time.perf_counter()
    c = ex.run_sweep(T, lists[0], thetas)
    t_sweep = time.perf_counter() - t0
    same = all(np.array_equal(x, y) for x, y in zip(a, b)) and all(np.array_equal(x, y) for x, y in zip(a, c))
    print(json.dumps({"config": name, "qubits": n, "circuits": len(lists), "trajectories_per_circuit": T, "seconds_loop": t_loop,
                      "seconds_run_many": t_many, "seconds_run_sweep": t_sweep, "circuits_per_s_loop": len(lists) / t_loop,
                      "circuits_per_s_run_many": len(lists) / t_many, "circuits_per_s_run_sweep": len(lists) / t_sweep,
                      "kernel_ms_per_circuit": ex.last_stats["kernel_ms"], "identical": same}), flush=True)


_pairs12 = circuits.coupled_pairs(h_tq, 12, adjacent_only=True)
sweep("C3 sweep: 20 QNN parameter sets x1e4", 12, [circuits.qnn_circuit(12, 4, _pairs12, seed=100 + i) for i in range(20)], h_sq, h_tq,
      2000 if quick else 10000)
