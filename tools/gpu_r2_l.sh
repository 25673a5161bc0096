#!/bin/bash
# round 2, GPU call L: the plain structure's two-launch LM step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
{ echo "== fast"; python tools/lm_timing.py 10000 0; echo "== general"; VG_LM_NOFAST=1 python tools/lm_timing.py 10000 0; echo "== fast MEI"; python tools/lm_timing.py 10000 2; echo "== general MEI"; VG_LM_NOFAST=1 python tools/lm_timing.py 10000 2; VG_LM_TRACE=1 python tools/lm_timing.py 10000 0 2>&1 | tail -8; } > gpurun_out/r2l_lm.txt 2>&1; cat gpurun_out/r2l_lm.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2l_lm_launches.csv python tools/lm_timing.py 10000 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2l_lm_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[10:20]:
    print(r[4][:50].ljust(50), r[7], r[8], r[-1])
PY
