#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in conv models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_$t.log | tail -n 8
done
for m in "retinanet 16" "yolov3 32" "fcos 4" "ssd300 64"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/ops_$n.txt)"
done
