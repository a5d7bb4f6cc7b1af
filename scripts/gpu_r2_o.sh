#!/bin/bash
# round 2, GPU call O: final evidence with the final code
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/o_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/o_pytest.log
tail -4 gpurun_out/o_pytest.log
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or replay_eagle or random_circuits or shared_no_jump or sharding or reordered_engine or philox or unitaries" ) > gpurun_out/o_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/o_pytest_spec.log
tail -3 gpurun_out/o_pytest_spec.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; tail -1 gpurun_out/o_smoke.log
python bench.py > gpurun_out/o_bench_default.json 2> gpurun_out/o_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/o_bench_default.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "strong", round(d["strong_scaling"]["value"]), d["clocks"], d["parity"]["max_abs_diff_probability"], d["c2_density_matrix"]["seconds"], d["cpu_baseline"]["value"])
PY
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/o_ref.json 2> gpurun_out/o_ref.err; cut -c1-200 gpurun_out/o_ref.json
python scripts/bench_configs.py > gpurun_out/o_configs.jsonl 2> gpurun_out/o_configs.err
cut -c1-230 gpurun_out/o_configs.jsonl
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/o_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/o_launches.csv $BENCH1 > gpurun_out/o_ncu_list.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+$' -s 30 -c 3 -o gpurun_out/o_prof_seg $BENCH1 > gpurun_out/o_ncu_seg.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+_ev' -s 60 -c 3 -o gpurun_out/o_prof_ev $BENCH1 > gpurun_out/o_ncu_ev.log 2>&1
ls -la gpurun_out/o_* | head -30
