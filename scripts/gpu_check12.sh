#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 -x -k "pool or resident" > gpurun_out/test_pool.log 2>&1
echo "pool tests exit $?"; grep -E "passed|failed|FAILED|Error|assert|rror" gpurun_out/test_pool.log | tail -n 20
for t in conv tail models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_$t.log | tail -n 14
done
for m in "ssd300 64" "retinanet 16" "yolov3 32"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$n.txt 2>&1; echo "== $m"; grep -E "CUDA-graph" gpurun_out/ops_$n.txt; head -12 gpurun_out/ops_$n.txt | cut -c1-100
  ODT_TC_WRES=0 timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_${n}_nowres.txt 2>&1; grep -E "CUDA-graph" gpurun_out/ops_${n}_nowres.txt
done
