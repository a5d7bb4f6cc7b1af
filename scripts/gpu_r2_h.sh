#!/bin/bash
# round 2, GPU call H: 5..8-qubit unitaries on the device, C3 through the tiled path
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "unitaries or pure_state or density_matrix_vs or replay_resident or c3_qnn" ) > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log
tail -5 gpurun_out/h_pytest.log
python scripts/exp_c3_tiled.py > gpurun_out/h_c3_tiled.jsonl 2> gpurun_out/h_c3_tiled.err
cat gpurun_out/h_c3_tiled.jsonl; tail -3 gpurun_out/h_c3_tiled.err
