#!/usr/bin/env python
"""Small runs of every generated kernel family for compute-sanitizer (racecheck / memcheck): resident (trunk + trajectories,
Heron and Eagle), tiled whole-segment + event kernels (literal and parametric), interpreter tile / resident kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from noisycircuits_b200 import circuits, noise_io, _lib
from noisycircuits_b200.program import PackedCircuit, Program
hs, ht, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
es, et, _m2, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "eagle16_noise.npz"))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = []
if which in ("all", "resident"):
    cases += [("resident heron n=9 generated", 9, circuits.heron_layered(9, 3, circuits.coupled_pairs(ht, 9), seed=1), hs, ht, dict(specialize="parametric"), 96),
              ("resident eagle n=8 generated", 8, circuits.eagle_layered(8, 2, circuits.coupled_pairs(et, 8), seed=2), es, et, dict(specialize="parametric"), 64),
              ("resident heron n=9 interpreter", 9, circuits.heron_layered(9, 3, circuits.coupled_pairs(ht, 9), seed=1), hs, ht, dict(), 64)]
if which in ("all", "tiled"):
    cases += [("tiled heron n=13 literal", 13, circuits.heron_layered(13, 2, circuits.coupled_pairs(ht, 13), seed=3), hs, ht, dict(specialize=True, tile_bits=10), 24),
              ("tiled eagle n=13 parametric", 13, circuits.eagle_layered(13, 1, circuits.coupled_pairs(et, 13), seed=4), es, et, dict(specialize="parametric", tile_bits=10), 16),
              ("tiled heron n=13 interpreter", 13, circuits.heron_layered(13, 2, circuits.coupled_pairs(ht, 13), seed=3), hs, ht, dict(tile_bits=10), 16)]
for name, n, instr, sq, tq, opts, T in cases:
    prog = Program(PackedCircuit(n, instr, sq, tq), _lib.MODE_MCWF, **opts)
    rng = np.random.default_rng(0)
    U = rng.random((T, len(instr)))
    U[T // 2:] = 1 - 0.05 * rng.random((T - T // 2, len(instr))) ** 2
    p, st = prog.run_mcwf(T, uniforms=U)
    q, st2 = prog.run_mcwf(T, seed=5)
    print(name, "norms", float(p.sum()) / T, float(q.sum()) / T, "events", st["events"], st["hard_events"], "spec", st["specialized_launches"], flush=True)
print("done")
