#!/bin/bash
# round 2, GPU call S: two-phase runs (launch / collect), pipelined sweep
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "generated_resident or two_phase or run_many or shared_jump_free or c3_qnn or sharding or execute_chain" ) > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log
tail -5 gpurun_out/s_pytest.log
timeout 600 python scripts/exp_c3_resident.py > gpurun_out/s_c3.jsonl 2> gpurun_out/s_c3.err
head -3 gpurun_out/s_c3.jsonl | cut -c1-450; tail -5 gpurun_out/s_c3.err
NCB_QUICK= timeout 300 python scripts/exp_determinism.py 2>&1 | tail -8
