#!/bin/bash
mkdir -p gpurun_out
VG_LM_NOFUSE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2k_lm_launches_nofuse.csv python tools/lm_timing.py 10000 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2k_lm_launches_nofuse.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[10:24]:
    print(r[4][:50].ljust(50), r[7], r[8], r[-1])
PY
