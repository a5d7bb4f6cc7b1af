#!/bin/bash
# round 2, GPU call P: generated resident kernels (states that fit one CTA): parity, C3 / C1 timings against the interpreting kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "generated_resident or replay_resident or shared_jump_free or c3_qnn or run_many" ) > gpurun_out/p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/p_pytest.log
tail -5 gpurun_out/p_pytest.log
timeout 900 python scripts/exp_c3_resident.py > gpurun_out/p_c3.jsonl 2> gpurun_out/p_c3.err
cat gpurun_out/p_c3.jsonl; tail -5 gpurun_out/p_c3.err
NCB_VERBOSE=1 timeout 300 python - > gpurun_out/p_verbose.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.solvers import RemoteExecutor
sq, tq, _m, _ = noise_io.load_noise_tables(os.path.join("tests", "golden", "heron20_noise.npz"))
pairs = circuits.coupled_pairs(tq, 12, adjacent_only=True)
lists = [circuits.qnn_circuit(12, 4, pairs, seed=100 + i) for i in range(8)]
ex = RemoteExecutor(12, sq, tq, 1)
out = ex.run_many(10000, lists)
print([float(o.sum()) for o in out], ex.last_stats)
PY
grep -c "structure hash" gpurun_out/p_verbose.log; grep "ncb\]" gpurun_out/p_verbose.log | sort | uniq -c | head -20; tail -2 gpurun_out/p_verbose.log
