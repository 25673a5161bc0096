set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in 0 1; do
VG_PDL=$v python tools/kernel_timing.py --modes full,normal 2>&1 | tail -2
VG_PDL=$v python tools/kernel_timing.py --n-img 25000 --modes full 2>&1 | tail -1
VG_PDL=$v python tools/kernel_timing.py --model 2 --modes full 2>&1 | tail -1
VG_PDL=$v python tools/lm_timing.py 2>&1 | tail -1
VG_PDL=$v python bench.py --cpu-seconds 0.5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['roofline']['frac'], d['e2e']['value'], d['lm']['iters_per_s'])"
done
