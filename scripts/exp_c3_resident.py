#!/usr/bin/env python
"""C3 (12-qubit QNN, 10^4 trajectories) and C1 (4 qubits, 1000): interpreting resident kernel against the generated resident
kernels (R = 3 / 4), and the parameter sweep (run_many / run_sweep) on top of them."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from noisycircuits_b200 import cache, circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))


def one(tag, n, instr, T, lists=None, env=None, **opts):
    for k in ("NCB_REG_BITS", "NCB_CTAS_PER_SM", "NCB_RES_FAST"):
        os.environ.pop(k, None)
    os.environ.update(env or {})
    cache.clear_programs()
    ex = RemoteExecutor(n, sq, tq, 1, **opts)
    t0 = time.perf_counter(); ex.run(T, instr); first = time.perf_counter() - t0
    best = None
    for _ in range(5):
        t0 = time.perf_counter(); p = ex.run(T, instr); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ex.last_stats
    out = {"variant": tag, "qubits": n, "trajectories": T, "first_call_s": first, "seconds": best, "kernel_ms": st["kernel_ms"],
           "spec_launches": st["specialized_launches"], "events": st["events"], "hard": st["hard_events"], "norm": float(p.sum())}
    if lists:
        ex.run_many(T, lists[:6])
        t0 = time.perf_counter(); a = ex.run_many(T, lists); out["run_many_circuits_per_s"] = len(lists) / (time.perf_counter() - t0)
        thetas = np.array([[0.0 if q is None else float(q) for _g, _q, q in l] for l in lists])
        t0 = time.perf_counter(); b = ex.run_sweep(T, lists[0], thetas); out["run_sweep_circuits_per_s"] = len(lists) / (time.perf_counter() - t0)
        out["sweep_identical"] = all(np.array_equal(x, y) for x, y in zip(a, b))
        out["sweep_kernel_ms"] = ex.last_stats["kernel_ms"]
    print(json.dumps(out), flush=True)
    return p


n, T = 12, 10000
pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
lists = [circuits.qnn_circuit(n, 4, pairs, seed=100 + i) for i in range(40)]
ref = one("C3 interpreter", n, lists[0], T, lists, specialize=False)
for tag, env in (("C3 generated", {}), ("C3 generated, no fast path", {"NCB_RES_FAST": "0"}), ("C3 generated again", {}), ("C3 generated 1 CTA/SM", {"NCB_CTAS_PER_SM": "1"})):
    p = one(tag, n, lists[0], T, lists, env=env, specialize="parametric")
    print(json.dumps({"variant": tag, "max_abs_diff_vs_interpreter": float(np.abs(p - ref).max())}), flush=True)
for T2 in (1000, 100000):
    one("C3 interpreter T=%d" % T2, n, lists[0], T2, specialize=False)
    one("C3 generated T=%d" % T2, n, lists[0], T2, specialize="parametric")
i4 = circuits.heron_ansatz(4, 4, circuits.coupled_pairs(tq, 4), seed=42)
one("C1 interpreter", 4, i4, 1000, specialize=False)
one("C1 generated", 4, i4, 1000, specialize="parametric")
i10 = circuits.heron_layered(10, 10, circuits.coupled_pairs(tq, 10), seed=42)
one("10q depth 10 interpreter", 10, i10, 20000, specialize=False)
one("10q depth 10 generated", 10, i10, 20000, specialize="parametric")
