#!/usr/bin/env python
"""C3 as a training loop writes it: RemoteExecutor.run once per parameter set (no sweep call)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisycircuits_b200 import cache, circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
pairs = circuits.coupled_pairs(tq, 12, adjacent_only=True)
lists = [circuits.qnn_circuit(12, 4, pairs, seed=100 + i) for i in range(48)]
ex = RemoteExecutor(12, sq, tq, 1)
for l in lists[:8]:
    ex.run(10000, l)
t0 = time.perf_counter()
for l in lists[8:]:
    p = ex.run(10000, l)
dt = time.perf_counter() - t0
print("plain loop of run(): %.1f circuits/s (%.2f ms per circuit)" % (40 / dt, dt / 40 * 1e3), cache.stats, ex.last_stats["specialized_launches"])
