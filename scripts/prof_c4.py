#!/usr/bin/env python
"""C4 (16-qubit Eagle circuit, depth 40) on the generated kernels, one call of 4736 trajectories: the command of the ncu launch list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "eagle16_noise.npz"))
n, T = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 4736
instr = circuits.eagle_layered(n, 40, circuits.coupled_pairs(tq, n), seed=42)
ex = RemoteExecutor(n, sq, tq, 1, specialize=True)
ex.run(64, instr)
t0 = time.perf_counter(); p = ex.run(T, instr); dt = time.perf_counter() - t0
print(float(p.sum()), T / dt, ex.last_stats)
