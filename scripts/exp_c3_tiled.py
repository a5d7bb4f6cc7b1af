#!/usr/bin/env python
"""C3 (12-qubit QNN, 10^4 trajectories): single-CTA resident kernel against the tiled path (tile_bits 11 / 10)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
n, T = 12, 10000
pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
lists = [circuits.qnn_circuit(n, 4, pairs, seed=100 + i) for i in range(12)]
ref = None
for name, opts in (("resident", {}), ("tiled K=11 interp", dict(tile_bits=11, specialize=False)), ("tiled K=11 parametric", dict(tile_bits=11, specialize="parametric")),
                   ("tiled K=10 parametric", dict(tile_bits=10, specialize="parametric")), ("tiled K=11 literal", dict(tile_bits=11, specialize=True))):
    ex = RemoteExecutor(n, sq, tq, 1, **opts)
    ex.run(T, lists[0])
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); p = ex.run(T, lists[0]); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ex.last_stats
    if ref is None:
        ref = p
    t0 = time.perf_counter(); ex.run_many(T, lists); t_many = time.perf_counter() - t0
    print(json.dumps({"variant": name, "seconds_per_circuit": best, "kernel_ms": st["kernel_ms"], "sweeps": st["sweeps"], "launches": st["launches"],
                      "max_abs_diff_vs_resident": float(np.abs(p - ref).max()), "run_many_circuits_per_s": len(lists) / t_many}), flush=True)
