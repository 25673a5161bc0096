#!/bin/bash
# compute-sanitizer over the new kernels (detector, point API) and one overlapping-launch burst
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_detector_gpu.py -x -q -m gpu \
    -k "grid_without or refined_corners or subpixel_cost or refinement_alone or other_board" > gpurun_out/san_mem_det.log 2>&1
echo "memcheck detector rc=$?"; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/san_mem_det.log; tail -3 gpurun_out/san_mem_det.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_camera_api.py tests/test_priors_gpu.py -x -q -m gpu \
    -k "odometry_cost or project or reconstruct or cloud" > gpurun_out/san_mem_api.log 2>&1
echo "memcheck api rc=$?"; tail -3 gpurun_out/san_mem_api.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_detector_gpu.py -x -q -m gpu \
    -k "refinement_alone or subpixel_cost" > gpurun_out/san_race_det.log 2>&1
echo "racecheck refine rc=$?"; tail -3 gpurun_out/san_race_det.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_solve_gpu.py -x -q -m gpu -k "overlapping and 2-1500" > gpurun_out/san_mem_overlap.log 2>&1
echo "memcheck overlap rc=$?"; tail -3 gpurun_out/san_mem_overlap.log
