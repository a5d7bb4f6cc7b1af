#!/bin/bash
# round 2, GPU call Y: last check of the final tree (full GPU suite, smoke, short default bench)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/y_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/y_pytest.log
tail -4 gpurun_out/y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
python bench.py --no-parity > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/y_bench.json"))
print("default", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), d["clocks"], d["specialize_build"], "c3", d["c3_qnn_sweep"]["trajectories_per_s"], d["c3_qnn_sweep"]["sweep"])
PY
tail -2 gpurun_out/y_bench.err
