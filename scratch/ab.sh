#!/bin/bash
# A/B: current libncb200.so vs scratch/libprev.bin, alternating, same box.  usage: scratch/ab.sh [bench args]
cp noisycircuits_b200/libncb200.so /tmp/new.so
for rep in 1 2; do
  for w in new prev; do
    if [ $w = new ]; then cp /tmp/new.so noisycircuits_b200/libncb200.so; else cp scratch/libprev.bin noisycircuits_b200/libncb200.so; fi
    python bench.py --steps 3 --warmup 2 --traj-per-step 296 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3))"
  done
done
cp /tmp/new.so noisycircuits_b200/libncb200.so
