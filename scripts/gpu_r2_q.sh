#!/bin/bash
# round 2, GPU call Q: ncu on the generated resident kernels (C3)
mkdir -p gpurun_out
python scripts/prof_c3.py > gpurun_out/q_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/q_launches.csv python scripts/prof_c3.py > gpurun_out/q_ncu_list.log 2>&1
tail -2 gpurun_out/q_plain.log
python scripts/prof_c3.py > gpurun_out/q_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ncb_res -s 4 -c 2 -o gpurun_out/q_prof_res python scripts/prof_c3.py > gpurun_out/q_ncu_res.log 2>&1
tail -3 gpurun_out/q_ncu_res.log
grep -v "^==" gpurun_out/q_launches.csv | python -c "
import csv, sys, collections
rows = list(csv.DictReader(sys.stdin))
agg = collections.OrderedDict()
for r in rows[-40:]:
    print(r['Kernel Name'][:60], r['Metric Value'], r.get('Grid Size'), r.get('Block Size'))
"
