#!/usr/bin/env python
"""Compiles the generated kernels of the benchmarked / smoke-tested programs into the in-tree cubin cache (host only, NVRTC):
the cache travels with the tree, so a GPU box starts without compiling.  Run after the library has been built."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from noisycircuits_b200 import circuits, noise_io, _lib
from noisycircuits_b200.program import PackedCircuit, Program
hs, ht, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
es, et, _m2, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "eagle16_noise.npz"))
sq, tq, _meas, _pairs, c5 = bench.workload()
jobs = [("C5 20q Heron depth 50, literal", bench.NQ, c5, sq, tq, dict(specialize=True)),
        ("C3 12q QNN, resident", 12, circuits.qnn_circuit(12, 4, circuits.coupled_pairs(ht, 12, adjacent_only=True), seed=100), hs, ht, dict(specialize="parametric")),
        ("C4 16q Eagle depth 40, literal", 16, circuits.eagle_layered(16, 40, circuits.coupled_pairs(et, 16), seed=42), es, et, dict(specialize=True)),
        ("smoke 14q literal", 14, circuits.heron_layered(14, 4, circuits.coupled_pairs(ht, 14), seed=3), hs, ht, dict(specialize=True)),
        ("smoke 10q resident", 10, circuits.qnn_circuit(10, 3, circuits.coupled_pairs(ht, 10, adjacent_only=True), seed=5), hs, ht, dict(specialize="parametric"))]
for name, n, instr, s1, s2, opts in jobs:
    t0 = time.time()
    ok = Program(PackedCircuit(n, instr, s1, s2), _lib.MODE_MCWF, **opts).specialize_build()
    print(f"{name}: {'built' if ok else 'FAILED'} in {time.time() - t0:.1f} s", flush=True)
