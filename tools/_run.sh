set -x
python -m pytest tests/test_priors_gpu.py -q --tb=short 2>&1 | tail -60
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_priors_gpu.py -x -q > gpurun_out/sanitizer_priors.log 2>&1; echo sanitizer rc=$?; tail -4 gpurun_out/sanitizer_priors.log
python -m pytest tests -m gpu -q 2>&1 | tail -3
