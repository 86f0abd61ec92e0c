#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r2c11
python tools/round2/launch_diag.py powerlaw > ${P}_diag_pl.txt 2>&1; cat ${P}_diag_pl.txt | tail -9
python tools/round2/launch_diag.py er > ${P}_diag_er.txt 2>&1; cat ${P}_diag_er.txt | tail -9
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:walk_uw_edge --csv --log-file ${P}_diag_launches.csv python tools/round2/launch_diag.py powerlaw > ${P}_diag_ncu.txt 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c11_diag_launches.csv')) if len(r)>5 and 'walk_uw_edge' in ''.join(r)]
for r in rows: print(r[-1], r[4] if len(r)>4 else '', [x for x in r if 'grid' in x.lower() or x.startswith('(')][:2])
PY
