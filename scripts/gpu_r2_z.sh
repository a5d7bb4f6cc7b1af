#!/bin/bash
# round 2, GPU call Z: event kernels whose whole passes take the fast lists (NCB_EV_FAST): parity, A/B on C5 and C4
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "specialised or generated_event or c5_full or c4_eagle or replay_tiled or replay_eagle or random_circuits or shared_no_jump or sharding" ) > gpurun_out/z_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/z_pytest.log
tail -3 gpurun_out/z_pytest.log
( time NCB_SPECIALIZE=1 timeout 600 python -m pytest tests -m gpu -x -q -k "replay_tiled_heron or replay_eagle or random_circuits or shared_no_jump or reordered_engine" ) > gpurun_out/z_pytest_spec.log 2>&1
echo "pytest exit $?" >> gpurun_out/z_pytest_spec.log
tail -3 gpurun_out/z_pytest_spec.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --no-c2 --strong-traj 0"
for v in 1 0 1 0; do
  NCB_EV_FAST=$v $B > gpurun_out/z_bench_$v.json 2> gpurun_out/z_bench_$v.err
  python -c "
import json; d = json.load(open('gpurun_out/z_bench_$v.json')); print('EV_FAST=$v value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), d['clocks']['sm_mhz'], d['specialize_build'])"
done
for v in 1 0; do NCB_EV_FAST=$v python scripts/prof_c4.py 2>&1 | tail -1 | cut -c1-60; done
