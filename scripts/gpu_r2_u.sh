#!/bin/bash
# round 2, GPU call U: generic one-bit products split into phase tables + rx-like operators (fast lists): parity, C4
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "eagle or c4 or random_circuits or replay_tiled or specialised or c3 or generated or c5_full or distribution or pure" ) > gpurun_out/u_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/u_pytest.log
tail -4 gpurun_out/u_pytest.log
( time NCB_SPECIALIZE=1 timeout 900 python -m pytest tests -m gpu -x -q -k "replay_eagle or c4_eagle or random_circuits" ) > gpurun_out/u_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/u_pytest_spec.log
tail -3 gpurun_out/u_pytest_spec.log
timeout 900 python scripts/exp_c4_split.py > gpurun_out/u_c4.jsonl 2> gpurun_out/u_c4.err
cat gpurun_out/u_c4.jsonl; tail -3 gpurun_out/u_c4.err
