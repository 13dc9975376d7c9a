#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 -k "stem" > gpurun_out/test_conv.log 2>&1; echo "stem tests exit $?"; tail -n 3 gpurun_out/test_conv.log
timeout 600 python scripts/profile_ops.py ssd300 64 2>&1 | grep -E "conv1_1|tail|decode|conv1_2"
timeout 600 python scripts/profile_ops.py retinanet 16 2>&1 | grep -E "tail|decode"
timeout 600 python scripts/profile_ops.py yolov3 32 2>&1 | tail -n 90 | cut -c1-118
timeout 600 python scripts/profile_ops.py fcos 4 2>&1 | grep -E "tail|decode|sum of"
