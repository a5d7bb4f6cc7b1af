#!/bin/bash
# 2-GPU variants: each arg = env assignments
for cfg in "$@"; do
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --steps 4 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg |', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1))"
done
