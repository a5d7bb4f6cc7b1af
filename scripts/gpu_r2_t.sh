#!/bin/bash
# round 2, GPU call T: launch list of C4 (16-qubit Eagle)
mkdir -p gpurun_out
python scripts/prof_c4.py > gpurun_out/t_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/t_launches.csv python scripts/prof_c4.py > gpurun_out/t_ncu_list.log 2>&1
tail -2 gpurun_out/t_plain.log
grep -v "^==" gpurun_out/t_launches.csv | python -c "
import csv, sys, collections, re
rows = list(csv.DictReader(sys.stdin))
agg = collections.OrderedDict()
for r in rows:
    k = re.sub(r'ncb_seg_\d+', 'ncb_seg_K', r['Kernel Name'])[:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value'])
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print('%-70s %5d launches %9.3f ms %5.1f%% avg %7.1f us' % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))
"
