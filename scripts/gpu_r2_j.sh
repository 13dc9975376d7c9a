#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/r2j_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "autograd loss" gpurun_out/r2j_gpu_tests.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2j_gpu_tests.log | tail -n 12
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool racecheck --racecheck-report all --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_tail.py -q -x -k "golden or live_oracle or spill or overflow" -p no:cacheprovider > gpurun_out/r2_sanitize_racecheck_tail_fixed.log 2>&1
echo "racecheck tail (after the fix): exit $? ; $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2_sanitize_racecheck_tail_fixed.log | tail -n 1) ; $(tail -n 1 gpurun_out/r2_sanitize_racecheck_tail_fixed.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
