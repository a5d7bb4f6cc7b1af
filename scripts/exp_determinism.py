#!/usr/bin/env python
"""Repeats one generated-resident run and compares the results bit for bit (single engine, then the sweep pipeline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from noisycircuits_b200 import circuits, noise_io, _lib
from noisycircuits_b200.program import PackedCircuit, Program
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "heron20_noise.npz"))
n, T = 12, 10000
pairs = circuits.coupled_pairs(tq, n, adjacent_only=True)
lists = [circuits.qnn_circuit(n, 4, pairs, seed=100 + i) for i in range(40)]
prog = Program(PackedCircuit(n, lists[0], sq, tq), _lib.MODE_MCWF, specialize="parametric")
ref, st = prog.run_mcwf(T)
bad = 0
for r in range(20):
    p, st2 = prog.run_mcwf(T)
    if not np.array_equal(p, ref):
        bad += 1
        print("single engine: run", r, "differs: max abs", float(np.abs(p - ref).max()), "events", st2["events"], st["events"], flush=True)
print("single engine repeats differing:", bad, "of 20")
# update path on one engine: alternate two circuits
pk = [PackedCircuit(n, l, sq, tq) for l in lists[:2]]
refs = []
for k in range(2):
    prog.update(pk[k]); refs.append(prog.run_mcwf(T)[0])
bad = 0
for r in range(400):
    prog.update(pk[r & 1]); p = prog.run_mcwf(T)[0]
    if not np.array_equal(p, refs[r & 1]):
        bad += 1
        if bad <= 3: print("update path: run", r, "differs: max abs", float(np.abs(p - refs[r & 1]).max()), flush=True)
print("update path repeats differing:", bad, "of 400")
if os.environ.get("NCB_QUICK"): sys.exit(0)
ex = RemoteExecutor(n, sq, tq, 1)
a = ex.run_many(T, lists)
for rep in range(3):
    b = ex.run_many(T, lists)
    d = [i for i in range(len(lists)) if not np.array_equal(a[i], b[i])]
    print("run_many repeat", rep, "differing circuits:", d, [float(np.abs(a[i] - b[i]).max()) for i in d], flush=True)
for threads in (1,):
    ex.PREPARE_THREADS = threads
    b = ex.run_many(T, lists)
    d = [i for i in range(len(lists)) if not np.array_equal(a[i], b[i])]
    print("run_many with", threads, "prepare thread: differing circuits:", d, [float(np.abs(a[i] - b[i]).max()) for i in d], flush=True)

# tiled generated kernels (whole-segment + event kernels, shared prefix, two streams): the same bits every time
nt, Tt = 15, 600
it = circuits.heron_layered(nt, 6, circuits.coupled_pairs(tq, nt), seed=9)
pt = Program(PackedCircuit(nt, it, sq, tq), _lib.MODE_MCWF, specialize=True)
ref_t, st_t = pt.run_mcwf(Tt, seed=2)
bad = sum(0 if np.array_equal(pt.run_mcwf(Tt, seed=2)[0], ref_t) else 1 for _ in range(30))
print("tiled generated kernels: repeats differing:", bad, "of 30; events", st_t["events"], "hard", st_t["hard_events"], "spec launches", st_t["specialized_launches"])
