#!/usr/bin/env python
"""Where does a backend.execute() call spend its time?  (diagnostic, GPU)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from noisycircuits_b200 import simulator, backend
from noisycircuits_b200.solvers import RemoteExecutor

sq, tq, meas, pairs, instr = bench.workload()
NQ = bench.NQ
solver = RemoteExecutor(NQ, sq, tq, 1, specialize=True)
tps = 592
for s in range(3):
    solver.run_sum(instr, s * tps, tps)
torch.cuda.synchronize()
meas_all = {q: meas[q] for q in range(NQ)}
for s in range(8):
    t0 = time.perf_counter()
    acc = simulator.DeviceBuffer(1 << NQ)
    t1 = time.perf_counter()
    solver.run_sum(instr, (10 + s) * tps, tps, device_probs=acc.ptr)
    t2 = time.perf_counter()
    out = simulator.execute_tail_device(acc.ptr, NQ, list(range(NQ)), meas_all, 1.0 / tps)
    t3 = time.perf_counter()
    acc.free()
    t4 = time.perf_counter()
    print("call %d: alloc %.2f ms  run_sum %.2f ms (kernel_ms %.2f)  tail %.2f ms  free %.2f ms" % (s, (t1 - t0) * 1e3, (t2 - t1) * 1e3, solver.last_stats["kernel_ms"], (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
for s in range(4):
    t0 = time.perf_counter()
    backend.execute(NQ, instr, sq, tq, meas_all, None, tps, executor=solver, traj_offset=(30 + s) * tps)
    print("execute() %d: %.2f ms" % (s, (time.perf_counter() - t0) * 1e3), flush=True)
for s in range(4):
    t0 = time.perf_counter()
    solver.run_sum(instr, (40 + s) * tps, tps)
    print("run_sum(host) %d: %.2f ms (kernel_ms %.2f)" % (s, (time.perf_counter() - t0) * 1e3, solver.last_stats["kernel_ms"]), flush=True)
