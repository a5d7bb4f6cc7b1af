#!/bin/bash
# round 2, 8-GPU box: the scaling curve on ONE box (N = 1, 2, 4, 8 as the driver launches the bench) and C4 as one sharded call
mkdir -p gpurun_out
OPTS="--steps 6 --warmup 4 --no-parity --no-c2 --no-cpu-baseline"
python bench.py --gpus 1 $OPTS > gpurun_out/x8_n1.json 2> gpurun_out/x8_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N $OPTS > gpurun_out/x8_n$N.json 2> gpurun_out/x8_n$N.err
done
for N in 1 2 4 8; do grep '^{' gpurun_out/x8_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'strong', round(d['strong_scaling']['value']), 'ms', round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"; done
python scripts/c4_sharded.py > gpurun_out/x8_c4_n1.json 2> gpurun_out/x8_c4_n1.err; cat gpurun_out/x8_c4_n1.json
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N scripts/c4_sharded.py > gpurun_out/x8_c4_n$N.json 2> gpurun_out/x8_c4_n$N.err; grep '^{' gpurun_out/x8_c4_n$N.json
done
tail -2 gpurun_out/x8_c4_n8.err
