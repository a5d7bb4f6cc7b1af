#!/bin/bash
# round 2, GPU call F: last-segment norm in the whole-segment kernel, event-launch grid, Eagle geometry, full suite + smoke
mkdir -p gpurun_out
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or random_circuits or shared_no_jump or sharding or philox or c5_full or execute" ) > gpurun_out/f_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest_spec.log
tail -4 gpurun_out/f_pytest_spec.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/f_bench_$name.json 2> gpurun_out/f_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/f_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run base
run evchunk4 NCB_EV_CHUNK=4
run evchunk8 NCB_EV_CHUNK=8
run evchunk16 NCB_EV_CHUNK=16
run batch592 NCB_BATCH=592 NCB_BATCH_MEM_GB=32
python - > gpurun_out/f_eagle.log 2>&1 <<'PY'
import os, sys, time, json
sys.path.insert(0, os.getcwd())
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
G = os.path.join("tests", "golden")
e_sq, e_tq, _em, _ = noise_io.load_noise_tables(os.path.join(G, "eagle16_noise.npz"))
instr = circuits.eagle_layered(16, 40, circuits.coupled_pairs(e_tq, 16), seed=42)
T = 4736
for name, opts in (("interp", dict(specialize=False)), ("literal_r4", dict(specialize=True)), ("literal_r3", dict(specialize=True, reg_bits=3)),
                   ("parametric_r4", dict(specialize="parametric")), ("parametric_r3", dict(specialize="parametric", reg_bits=3))):
    ex = RemoteExecutor(16, e_sq, e_tq, 1, **opts)
    ex.run(64, instr)
    best = None
    for _ in range(2):
        t0 = time.perf_counter(); p = ex.run(T, instr); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ex.last_stats
    print(json.dumps({"variant": name, "traj_per_s": T / best, "kernel_ms": st["kernel_ms"], "sweeps_per_traj": st["sweeps"] / T, "spec": st["specialized_launches"], "tile": st["tile_launches"]}), flush=True)
PY
cat gpurun_out/f_eagle.log
