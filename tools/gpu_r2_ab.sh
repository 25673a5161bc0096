#!/bin/bash
python tools/e2e_probe.py 2>&1 | sed -n 2,2p
nvidia-smi -i 0 --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 20 > /dev/null 2>&1 &
SMI=$!
sleep 1
python tools/e2e_probe.py 2>&1 | sed -n 2,2p
kill $SMI
nvidia-smi -i 0 --query-gpu=clocks.sm --format=csv,noheader,nounits -lms 200 > /dev/null 2>&1 &
SMI=$!
sleep 1
python tools/e2e_probe.py 2>&1 | sed -n 2,2p
kill $SMI
