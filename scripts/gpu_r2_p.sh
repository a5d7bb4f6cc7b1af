#!/bin/bash
# round 2, GPU call P: generated resident kernels (states that fit one CTA): parity, C3 / C1 timings against the interpreting kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "generated_resident or replay_resident or shared_jump_free or c3_qnn or run_many" ) > gpurun_out/p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/p_pytest.log
tail -5 gpurun_out/p_pytest.log
timeout 900 python scripts/exp_c3_resident.py > gpurun_out/p_c3.jsonl 2> gpurun_out/p_c3.err
cat gpurun_out/p_c3.jsonl; tail -5 gpurun_out/p_c3.err
