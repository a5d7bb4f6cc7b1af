#!/bin/bash
# round 2, GPU call G: final evidence -- whole GPU suite, smoke, default bench, configs, ncu launch list + full captures
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -4 gpurun_out/g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; tail -1 gpurun_out/g_smoke.log
python bench.py > gpurun_out/g_bench_default.json 2> gpurun_out/g_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/g_bench_default.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "strong", round(d["strong_scaling"]["value"]), d["clocks"], d["parity"], d["c2_density_matrix"]["seconds"])
PY
python scripts/bench_configs.py > gpurun_out/g_configs.jsonl 2> gpurun_out/g_configs.err
cut -c1-260 gpurun_out/g_configs.jsonl
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/g_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/g_launches.csv $BENCH1 > gpurun_out/g_ncu_list.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+$' -s 30 -c 3 -o gpurun_out/g_prof_seg $BENCH1 > gpurun_out/g_ncu_seg.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+_ev' -s 60 -c 3 -o gpurun_out/g_prof_ev $BENCH1 > gpurun_out/g_ncu_ev.log 2>&1
python scripts/prof_dm.py > gpurun_out/g_dm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 33 -c 4 -o gpurun_out/g_prof_dm python scripts/prof_dm.py > gpurun_out/g_ncu_dm.log 2>&1
ls -la gpurun_out/g_*
