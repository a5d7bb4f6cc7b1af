#!/bin/bash
# round 2, GPU call D: R = 4 as the default of the generated kernels -- full GPU suite, bench, e2e diagnostic, ncu
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
python scripts/diag_e2e.py > gpurun_out/d_diag_e2e.log 2>&1
tail -18 gpurun_out/d_diag_e2e.log
python bench.py > gpurun_out/d_bench_default.json 2> gpurun_out/d_bench_default.err
cut -c1-200 gpurun_out/d_bench_default.json
run() {  # name, env...
  name=$1; shift
  env "$@" NCB_BENCH_PREWARM=3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0 > gpurun_out/d_bench_$name.json 2> gpurun_out/d_bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/d_bench_$name.json"))
    print("$name", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), round(d["roofline"]["sweeps_per_trajectory"], 1), d["clocks"]["sm_mhz"], d["specialized_launches"])
except Exception as e:
    print("$name", "FAILED", e)
PY
}
run r4_ops64 NCB_MAX_SEG_OPS=64
run r4_ops40 NCB_MAX_SEG_OPS=40
run r4_chunk8 NCB_DIRECT_CHUNK=8
run r4_chunk32 NCB_DIRECT_CHUNK=32
run r4_low3 NCB_LOW_BITS=3
run r4_nooverlap NCB_OVERLAP=0
# ---- ncu: launch list of one bench step, then full captures of the three kernel families ----
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/d_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/d_launches.csv $BENCH1 > gpurun_out/d_ncu_list.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+$' -s 30 -c 3 -o gpurun_out/d_prof_seg $BENCH1 > gpurun_out/d_ncu_seg.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+_ev' -s 60 -c 3 -o gpurun_out/d_prof_ev $BENCH1 > gpurun_out/d_ncu_ev.log 2>&1
python scripts/prof_dm.py > gpurun_out/d_dm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 31 -c 4 -o gpurun_out/d_prof_dm python scripts/prof_dm.py > gpurun_out/d_ncu_dm.log 2>&1
ls -la gpurun_out/d_*
