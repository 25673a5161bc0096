python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/lm_timing.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lm.csv python tools/lm_timing.py 10000 6 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_lm.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:70],[]).append(float(r[vi]))
for k,v in agg.items(): print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v)/1e3:8.2f} us")
PY
