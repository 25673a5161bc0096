#!/bin/bash
python tools/lm_timing.py 2>&1 | tail -3
VG_LM_BLIND=200 python tools/lm_timing.py 2>&1 | grep -i blind
VG_LM_TRACE=1 python tools/lm_timing.py 2>&1 | tail -12
