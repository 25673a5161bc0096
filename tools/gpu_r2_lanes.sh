#!/bin/bash
for l in 2 3 4 6; do for c in 4 8; do echo "== lanes $l chunk $c"; VG_DETECT_LANES=$l VG_DETECT_CHUNK=$c python tools/detector_timing.py 128 2>&1 | grep "improve=1" ; done; done
