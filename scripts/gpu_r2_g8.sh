#!/bin/bash
# round 2, 8-GPU call: the bench exactly as the driver launches it at N = 8 (weak value, execute_sharded e2e, strong scaling)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
$TR bench.py --gpus 8 --steps 6 --warmup 4 > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err
echo "exit $?"; grep '^{' gpurun_out/g8_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('N=8 value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3),'strong',d['strong_scaling'],'clocks',d['clocks'],'spec build',d['specialize_build'])"
tail -3 gpurun_out/g8_bench.err
