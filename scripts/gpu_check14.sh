#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 300 -x > gpurun_out/test_models.log 2>&1
echo "test_gpu_models exit $?"; grep -E "passed|failed|FAILED|Error|assert|rror" gpurun_out/test_models.log | tail -n 14
for m in "ssd300 64" "retinanet 16" "yolov3 32" "fcos 4" "ssd300 1"; do
  n=$(echo $m | tr ' ' '_')
  for k in 1 8 12; do
    ODT_STREAMS=$k timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_${n}_s$k.txt 2>&1; echo "== $m streams=$k: $(grep -E 'CUDA-graph' gpurun_out/ops_${n}_s$k.txt)"
  done
done
