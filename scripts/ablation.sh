#!/bin/bash
# Throughput of the headline workload with one optimisation switched off at a time.
# usage (on a B200): scripts/ablation.sh > profiles/rNN_ablation.jsonl
for cfg in "" "NCB_SPECIALIZE=0" "NCB_SHARE_PREFIX=0" "NCB_DIRECT=0" "NCB_OVERLAP=0" "NCB_DIRECT_CHUNK=0" "NCB_FUSE_ROT=0" "NCB_PASS_ORDER=0" "NCB_REORDER=0" "NCB_MAX_SEG_OPS=20" "NCB_TILE_BITS=12" "NCB_LOW_BITS=3" "NCB_FUSE=0"; do
  steps=3; tps=592
  if [ "$cfg" = "NCB_FUSE=0" ]; then steps=1; tps=148; fi
  env $cfg python bench.py --no-cpu-baseline --steps $steps --warmup 2 --traj-per-step $tps 2>/dev/null | tail -1 | CFG="$cfg" python -c "
import sys, json, os
d = json.loads(sys.stdin.read())
print(json.dumps({'switch': os.environ['CFG'] or 'defaults', 'trajectories_per_s': round(d['value'], 1), 'e2e': round(d['e2e']['value'], 1),
                  'roofline_frac': round(d['roofline']['frac'], 3), 'sweeps_per_trajectory': round(d['roofline']['sweeps_per_trajectory'], 1),
                  'sm_mhz': d['clocks']['sm_mhz']}))"
done
