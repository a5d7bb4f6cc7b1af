#!/usr/bin/env python
"""C4 (16-qubit Eagle, depth 40): generic one-bit products split into phase tables + an rx-like operator (NCB_SPLIT_U) on / off,
R = 3 / 4."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noisycircuits_b200 import cache, circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "eagle16_noise.npz"))
n, T = 16, 4736
instr = circuits.eagle_layered(n, 40, circuits.coupled_pairs(tq, n), seed=42)
ref = None
for name, env in (("split R=3", {}), ("no split R=3", {"NCB_SPLIT_U": "0"}), ("split R=4", {"NCB_REG_BITS": "4"}), ("split R=3 again", {}),
                  ("split R=3 parametric", {"NCB_SPECIALIZE": "2"}), ("split interpreter", {"NCB_SPECIALIZE": "0"}), ("no split interpreter", {"NCB_SPECIALIZE": "0", "NCB_SPLIT_U": "0"})):
    for k in ("NCB_SPLIT_U", "NCB_REG_BITS", "NCB_SPECIALIZE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cache.clear_programs()
    ex = RemoteExecutor(n, sq, tq, 1, specialize=True)
    ex.run(64, instr)
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); p = ex.run(T, instr); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if ref is None:
        ref = p
    st = ex.last_stats
    print(json.dumps({"variant": name, "traj_per_s": T / best, "kernel_ms": st["kernel_ms"], "sweeps_per_traj": st["sweeps"] / T, "spec": st["specialized_launches"],
                      "tile": st["tile_launches"], "norm": float(p.sum()), "max_abs_diff_vs_first": float(abs(p - ref).max())}), flush=True)
