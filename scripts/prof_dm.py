#!/usr/bin/env python
"""BASELINE configs[1] for the profiler: 12-qubit density-matrix run (4^12 vectorised state) of a depth-10 layered Heron
circuit, twice (the second is the one to capture)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import DensityMatrixSolver
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
n = 12
instr = circuits.heron_layered(n, 10, circuits.coupled_pairs(tq, n), seed=42)
dm = DensityMatrixSolver(n, sq, tq, instr, 1)
for _ in range(2):
    p = dm.solve(list(range(n)))
print(dm._program.describe().splitlines()[0])
print("trace", p.sum())
