#!/bin/bash
# round 2, GPU call W: final evidence with the final code (tests, smoke, bench arms, configurations, ncu launch lists and captures)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/w_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/w_pytest.log
tail -4 gpurun_out/w_pytest.log
( time NCB_SPECIALIZE=1 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or replay_eagle or random_circuits or shared_no_jump or sharding or reordered_engine or philox or unitaries" ) > gpurun_out/w_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/w_pytest_spec.log
tail -3 gpurun_out/w_pytest_spec.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1; tail -1 gpurun_out/w_smoke.log
python bench.py > gpurun_out/w_bench_default.json 2> gpurun_out/w_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/w_bench_default.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "strong", round(d["strong_scaling"]["value"]), d["clocks"], d["parity"]["max_abs_diff_probability"], d["c2_density_matrix"]["seconds"], d["cpu_baseline"]["value"])
print("c3", d["c3_qnn_sweep"])
PY
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/w_ref.json 2> gpurun_out/w_ref.err; cut -c1-200 gpurun_out/w_ref.json
python scripts/bench_configs.py > gpurun_out/w_configs.jsonl 2> gpurun_out/w_configs.err
cut -c1-260 gpurun_out/w_configs.jsonl
python scripts/exp_determinism.py > gpurun_out/w_determinism.log 2>&1; tail -8 gpurun_out/w_determinism.log
BENCH1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none $BENCH1 > gpurun_out/w_plain1.log 2>&1 &&
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/w_launches.csv $BENCH1 > gpurun_out/w_ncu_list.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+$' -s 30 -c 3 -o gpurun_out/w_prof_seg $BENCH1 > gpurun_out/w_ncu_seg.log 2>&1
NCB_BENCH_PREWARM=0 NCB_BENCH_CLOCKS=none ncu --set full --clock-control none --import-source on -k 'regex:ncb_seg_[0-9]+_ev' -s 60 -c 3 -o gpurun_out/w_prof_ev $BENCH1 > gpurun_out/w_ncu_ev.log 2>&1
python scripts/prof_c3.py > gpurun_out/w_c3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/w_c3_launches.csv python scripts/prof_c3.py > gpurun_out/w_c3_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ncb_res$' -s 3 -c 1 -o gpurun_out/w_prof_res python scripts/prof_c3.py > gpurun_out/w_ncu_res.log 2>&1
python scripts/prof_c4.py > gpurun_out/w_c4_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/w_c4_launches.csv python scripts/prof_c4.py > gpurun_out/w_c4_ncu_list.log 2>&1
tail -1 gpurun_out/w_c3_plain.log; tail -1 gpurun_out/w_c4_plain.log
ls -la gpurun_out/w_* | head -40
