#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1200 python -m pytest tests/test_gpu_models.py -q --timeout 600 -s > gpurun_out/r2v_models.log 2>&1
echo "model tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2v_models.log | tail -n 6
