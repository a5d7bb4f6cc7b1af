#!/bin/bash
# round 2, GPU call N: complex scalar factored out of generic 2x2 operators (Eagle): parity with the generated kernels forced, C4
mkdir -p gpurun_out
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "c4_eagle or replay_eagle or random_circuits or replay_tiled_heron or c5_full or single_inline or specialised" ) > gpurun_out/n_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/n_pytest_spec.log
tail -3 gpurun_out/n_pytest_spec.log
python - > gpurun_out/n_eagle.log 2>&1 <<'PY'
import os, sys, time, json
sys.path.insert(0, os.getcwd())
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
G = os.path.join("tests", "golden")
e_sq, e_tq, _em, _ = noise_io.load_noise_tables(os.path.join(G, "eagle16_noise.npz"))
instr = circuits.eagle_layered(16, 40, circuits.coupled_pairs(e_tq, 16), seed=42)
T = 4736
for name, env in (("literal_r3_factor", {}), ("literal_r3_nofactor", {"NCB_SPEC_FACTOR": "0"}), ("literal_r3_factor_b", {}), ("literal_r4_factor", {"NCB_REG_BITS": "4"})):
    os.environ.pop("NCB_SPEC_FACTOR", None); os.environ.pop("NCB_REG_BITS", None)
    os.environ.update(env)
    ex = RemoteExecutor(16, e_sq, e_tq, 1, specialize=True)
    ex.run(64, instr)
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); p = ex.run(T, instr); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ex.last_stats
    from noisycircuits_b200 import cache; cache.clear_programs()
    print(json.dumps({"variant": name, "traj_per_s": T / best, "kernel_ms": st["kernel_ms"], "sweeps_per_traj": st["sweeps"] / T, "spec": st["specialized_launches"], "tile": st["tile_launches"], "norm": float(p.sum())}), flush=True)
PY
cat gpurun_out/n_eagle.log
