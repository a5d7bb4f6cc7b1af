#!/usr/bin/env python
"""C3 (12-qubit QNN circuit, 10^4 trajectories) on the generated resident kernels, a few calls: the command the ncu captures
of ncb_res_trunk / ncb_res are taken from."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
n, T = 12, int(sys.argv[1]) if len(sys.argv) > 1 else 10000
instr = circuits.qnn_circuit(n, 4, circuits.coupled_pairs(tq, n, adjacent_only=True), seed=100)
ex = RemoteExecutor(n, sq, tq, 1, specialize="parametric")
for _ in range(4):
    p = ex.run(T, instr)
print(float(p.sum()), ex.last_stats)
