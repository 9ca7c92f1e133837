#!/bin/bash
# measurement aid: per-kernel device time of one pass of the bench workload for the groups in $PGROUPS (ncu, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/kt.csv python tools/probe_groups.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.DictReader(l for l in open('gpurun_out/kt.csv') if l.startswith('"')))
agg=collections.defaultdict(float)
for r in rows: agg[r['Kernel Name'].split('(')[0].split('::')[-1][:40]]+=float(r['Metric Value'].replace(',',''))/2e6
print({k:round(v,2) for k,v in agg.items() if 'synth' not in k})
PY
