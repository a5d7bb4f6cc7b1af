#!/bin/bash
for cfg in "$@"; do
  envs="${cfg%%--*}"; args="${cfg#*--}"
  env $envs python bench.py --no-cpu-baseline --steps 3 --warmup 2 $args 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg |', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],1), round(d['roofline']['frac'],3), round(d['roofline']['sweeps_per_trajectory'],1))"
done
