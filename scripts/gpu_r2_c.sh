#!/bin/bash
# round 2, GPU call C: literal coefficients again, generated event kernels, tile geometries (R = 4, K = 12), new bench line
mkdir -p gpurun_out
( time NCB_SPECIALIZE=1 NCB_REG_BITS=4 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or random_circuits or specialised or shared_no_jump or sharding or reordered or philox" ) > gpurun_out/c_pytest_r4.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest_r4.log
tail -4 gpurun_out/c_pytest_r4.log
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialize_build"], d["specialized_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run r3k11 NCB_VERBOSE=1
run r3k11_ev0 NCB_SPEC_EVENTS=0
run r3k11_nbuf1 NCB_SPEC_NBUF=1
run r4k11 NCB_REG_BITS=4
run r4k11_nbuf1 NCB_REG_BITS=4 NCB_SPEC_NBUF=1
run r4k12 NCB_REG_BITS=4 NCB_TILE_BITS=12 NCB_MAX_SEG_OPS=120
run r3k12 NCB_TILE_BITS=12 NCB_MAX_SEG_OPS=120
run r3k11_ops30 NCB_MAX_SEG_OPS=30
python bench.py > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err
cut -c1-600 gpurun_out/c_bench_default.json
python bench.py --impl reference --full-circuit --steps 1 --warmup 0 > gpurun_out/c_ref_full.json 2> gpurun_out/c_ref_full.err
python bench.py --impl reference --steps 13 --warmup 0 > gpurun_out/c_ref_slices.json 2> gpurun_out/c_ref_slices.err
cut -c1-300 gpurun_out/c_ref_full.json; cut -c1-300 gpurun_out/c_ref_slices.json
