set -x
ncu --set full --clock-control none --import-source on -k regex:"pose_factor|pose_backsub" -s 6 -c 2 -o gpurun_out/prof_lm_r1i python tools/lm_timing.py > gpurun_out/ncu_lm.log 2>&1
tail -2 gpurun_out/ncu_lm.log
