#!/bin/bash
# round 2, GPU call L: resident CTAs per SM of the generated kernels (launch bounds), same box
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/l_bench_$name.json 2> gpurun_out/l_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/l_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run ctas4
run ctas5 NCB_SPEC_CTAS=5
run ctas3 NCB_SPEC_CTAS=3
run ctas4b
run ctas5b NCB_SPEC_CTAS=5
run ctas6 NCB_SPEC_CTAS=6
