#!/bin/bash
# round 2, GPU call J: cheaper fused decision (one fence), merged groups kept whole around in-line jumps (WF_PULL)
mkdir -p gpurun_out
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "single_inline or replay_tiled_heron or random_circuits or shared_no_jump or sharding or philox or c5_full or execute or specialised or reordered_engine" ) > gpurun_out/j_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest_spec.log
tail -4 gpurun_out/j_pytest_spec.log
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/j_bench_$name.json 2> gpurun_out/j_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/j_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run pull
run nopull NCB_EV_PULL=0
run pull2
run nopull2 NCB_EV_PULL=0
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/j_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/j_launches.csv $BENCH1 > gpurun_out/j_ncu_list.log 2>&1
