#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for m in "ssd300 64" "retinanet 16" "yolov3 32"; do timeout 300 python scripts/tail_micro.py $m 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_tail.py -q -m gpu --timeout 300 > gpurun_out/test_tail.log 2>&1; echo "tail tests exit $?"; tail -n 2 gpurun_out/test_tail.log
timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/ops_ssd300_64.txt 2>&1; grep -E "CUDA-graph" gpurun_out/ops_ssd300_64.txt; sed -n 3p gpurun_out/ops_ssd300_64.txt | cut -c1-100
