#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 300 -k "loss" -s > gpurun_out/test_loss.log 2>&1
echo "loss tests exit $?"; grep -E "image |passed|failed|FAILED|Error|assert" gpurun_out/test_loss.log | tail -n 20
timeout 900 python -m pytest tests/test_gpu_tail.py -q -m gpu --timeout 300 > gpurun_out/test_tail.log 2>&1
echo "tail tests exit $?"; tail -n 2 gpurun_out/test_tail.log
