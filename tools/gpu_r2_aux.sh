#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/aux_timing.py > gpurun_out/aux_timing.txt 2>&1
cat gpurun_out/aux_timing.txt | tail -14
