#!/bin/bash
# round 2, last GPU call: the final tree -- GPU suite, smoke, default bench line, launch list and event-kernel capture
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -4 gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/f_bench_default.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["clocks"], d["parity"]["max_abs_diff_probability"], d["specialize_build"], "c3", round(d["c3_qnn_sweep"]["trajectories_per_s"]), d["c3_qnn_sweep"]["sweep"]["circuits_per_s"])
PY
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/f_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/f_launches.csv $BENCH1 > gpurun_out/f_ncu_list.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+_ev' -s 60 -c 3 -o gpurun_out/f_prof_ev $BENCH1 > gpurun_out/f_ncu_ev.log 2>&1
ls -la gpurun_out/f_* | head
