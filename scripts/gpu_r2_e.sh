#!/bin/bash
# round 2, GPU call E: x-like rotations, cold-code hints in the event kernels, persistent scratch; bulk-copy evaluation; configs
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "specialised or c5_full or random_circuits or replay_tiled or reference_class or execute or run_many" ) > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
python bench.py --no-parity > gpurun_out/e_bench_default.json 2> gpurun_out/e_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/e_bench_default.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "strong", round(d["strong_scaling"]["value"]), d["clocks"])
PY
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/e_bench_$name.json 2> gpurun_out/e_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/e_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run chunk4 NCB_DIRECT_CHUNK=4
run chunk12 NCB_DIRECT_CHUNK=12
run k12 NCB_TILE_BITS=12 NCB_MAX_SEG_OPS=120
run parametric NCB_SPECIALIZE=2
run interp NCB_SPECIALIZE=0
./scripts/tma_eval/tile_copy_bench 128 5 > gpurun_out/e_tile_copy_bench.json 2> gpurun_out/e_tile_copy_bench.err
cat gpurun_out/e_tile_copy_bench.json | cut -c1-220
python scripts/bench_configs.py > gpurun_out/e_configs.jsonl 2> gpurun_out/e_configs.err
cut -c1-330 gpurun_out/e_configs.jsonl
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/e_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/e_launches.csv $BENCH1 > gpurun_out/e_ncu_list.log 2>&1
