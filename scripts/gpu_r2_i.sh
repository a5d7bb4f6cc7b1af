#!/bin/bash
# round 2, GPU call I: branch decision fused into the generated event kernels -- parity, bench, launch list
mkdir -p gpurun_out
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or random_circuits or shared_no_jump or sharding or philox or c5_full or execute or specialised or reordered_engine or unitaries" ) > gpurun_out/i_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest_spec.log
tail -4 gpurun_out/i_pytest_spec.log
python bench.py --no-parity > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/i_bench_default.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "strong", round(d["strong_scaling"]["value"]), d["clocks"], d["gpu_launches"])
PY
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/i_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/i_launches.csv $BENCH1 > gpurun_out/i_ncu_list.log 2>&1
