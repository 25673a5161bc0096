set -x
tools/_bin/fp64_probe
VG_VARIANT=phase python tools/phase_clocks.py 10000 full
VG_VARIANT=phase python tools/phase_clocks.py 10000 normal
