#!/bin/bash
# round 2, GPU call R: whole GPU suite after the cfma change; resident kernels with the warp-per-trajectory planner; all configurations
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r_pytest.log
tail -5 gpurun_out/r_pytest.log
timeout 600 python scripts/exp_c3_resident.py > gpurun_out/r_c3.jsonl 2> gpurun_out/r_c3.err
cut -c1-420 gpurun_out/r_c3.jsonl; tail -5 gpurun_out/r_c3.err
timeout 900 python scripts/bench_configs.py > gpurun_out/r_configs.jsonl 2> gpurun_out/r_configs.err
cut -c1-330 gpurun_out/r_configs.jsonl; tail -3 gpurun_out/r_configs.err
