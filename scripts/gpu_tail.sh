#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for k in 4 8; do
for m in "ssd300 64" "retinanet 16" "yolov3 32"; do ODT_DECODE_BLOCKS_PER_SM=$k timeout 300 python scripts/tail_micro.py $m 2>&1 | tail -1 | sed "s/^/blocks_per_sm=$k /"; done; done
timeout 600 python -m pytest tests/test_gpu_tail.py -q -m gpu --timeout 300 > gpurun_out/test_tail.log 2>&1; echo "tail tests exit $?"; tail -n 2 gpurun_out/test_tail.log
