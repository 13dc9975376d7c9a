#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 -k "stem" > gpurun_out/test_stem.log 2>&1
echo "stem tests exit $?"; grep -E "passed|failed|FAILED|Error|assert|rror" gpurun_out/test_stem.log | tail -n 8
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --timeout 300 > gpurun_out/test_models.log 2>&1
echo "test_gpu_models exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_models.log | tail -n 8
for m in "retinanet 16" "fcos 4" "ssd300 64" "yolov3 32"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/ops_$n.txt)"; sed -n 2p gpurun_out/ops_$n.txt | cut -c1-100
done
