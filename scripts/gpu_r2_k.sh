#!/bin/bash
# round 2, GPU call K: decide launch vs decision in the event kernel (unrolled sum), same box
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/k_bench_$name.json 2> gpurun_out/k_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/k_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run separate
run fused NCB_FUSED_DECIDE=1
run separate2
run fused2 NCB_FUSED_DECIDE=1
run separate_nopull NCB_EV_PULL=0
( time NCB_SPECIALIZE=1 NCB_FUSED_DECIDE=1 python -m pytest tests -m gpu -x -q -k "single_inline or replay_tiled_heron or random_circuits or philox" ) > gpurun_out/k_pytest_fused.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest_fused.log
tail -3 gpurun_out/k_pytest_fused.log
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/k_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/k_launches.csv $BENCH1 > gpurun_out/k_ncu_list.log 2>&1
