#!/bin/bash
# round 2, GPU call M: real scalars factored out of the rotations (half their FP64 instructions): parity + A/B on one box
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "specialised or c5_full or single_inline or execute or smoke or philox" ) > gpurun_out/m_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest.log
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or random_circuits or shared_no_jump or sharding or reordered_engine" ) > gpurun_out/m_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/m_pytest_spec.log
tail -3 gpurun_out/m_pytest.log; tail -3 gpurun_out/m_pytest_spec.log
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/m_bench_$name.json 2> gpurun_out/m_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/m_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"], d["gpu_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run factor
run nofactor NCB_SPEC_FACTOR=0
run factor2
run nofactor2 NCB_SPEC_FACTOR=0
run factor_k12 NCB_TILE_BITS=12 NCB_MAX_SEG_OPS=120
python scripts/bench_configs.py > gpurun_out/m_configs.jsonl 2> gpurun_out/m_configs.err
cut -c1-250 gpurun_out/m_configs.jsonl
