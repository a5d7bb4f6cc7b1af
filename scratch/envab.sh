#!/bin/bash
# usage: scratch/envab.sh "ENV1=.. ENV2=.." "ENV.." ...   (each arg = one configuration; alternates twice)
for rep in 1 2; do
  for cfg in "$@"; do
    env $cfg python bench.py --steps 3 --warmup 2 --traj-per-step 296 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3))"
  done
done
