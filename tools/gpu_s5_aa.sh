#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/config_bench.py neg_shared skip-gram 512 128 37 500 adam 2>&1 | grep -v Warn
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/s5aa_adam_launches.csv python tools/config_bench.py neg_shared skip-gram 512 128 37 40 adam > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/s5aa_adam_launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split("(")[0][:60]].append(float(r[-1]))
for k,v in agg.items(): print("%-62s n=%3d avg %.2f us"%(k,len(v),sum(v)/len(v)/ (1000 if max(v)>1000 else 1)))
PY
