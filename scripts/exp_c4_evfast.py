import os, sys, time, json
sys.path.insert(0, os.getcwd())
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join("tests", "golden", "eagle16_noise.npz"))
instr = circuits.eagle_layered(16, 40, circuits.coupled_pairs(tq, 16), seed=42)
ex = RemoteExecutor(16, sq, tq, 1, specialize=True)
ex.run(64, instr)
for _ in range(3):
    t0 = time.perf_counter(); p = ex.run(4736, instr); dt = time.perf_counter() - t0
    print(os.environ.get("NCB_EV_FAST"), round(4736 / dt), "traj/s kernel_ms", round(ex.last_stats["kernel_ms"], 1), "spec", ex.last_stats["specialized_launches"], "tile", ex.last_stats["tile_launches"], flush=True)
