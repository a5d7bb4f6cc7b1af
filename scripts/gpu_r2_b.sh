#!/bin/bash
# round 2, GPU call B: generated kernels v2 (constant-bank coefficients, first pass loaded into registers, event-piece
# kernels) -- the GPU suite with every lean tiled program forced through them, then the bench per variant
mkdir -p gpurun_out
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "not c2_density and not c4_eagle" ) > gpurun_out/b_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest_spec.log
tail -4 gpurun_out/b_pytest_spec.log
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_$name.json 2> gpurun_out/b_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), d["clocks"]["sm_mhz"], d["specialize_build"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run l0_ev0 NCB_SPEC_LOAD=0 NCB_SPEC_EVENTS=0
run l0_ev1 NCB_SPEC_LOAD=0 NCB_SPEC_EVENTS=1
run l1_ev1 NCB_SPEC_LOAD=1 NCB_SPEC_EVENTS=1
run l2_ev1 NCB_SPEC_LOAD=2 NCB_SPEC_EVENTS=1
run l1_ev1_nbuf1 NCB_SPEC_LOAD=1 NCB_SPEC_EVENTS=1 NCB_SPEC_NBUF=1
run l2_ev1_nbuf1 NCB_SPEC_LOAD=2 NCB_SPEC_EVENTS=1 NCB_SPEC_NBUF=1
run l1_ev1_chunk0 NCB_SPEC_LOAD=1 NCB_SPEC_EVENTS=1 NCB_DIRECT_CHUNK=0
