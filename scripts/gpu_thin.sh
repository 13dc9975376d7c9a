#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 > gpurun_out/test_conv.log 2>&1
echo "test_gpu_conv exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_conv.log | tail -n 8
for m in "ssd300 64" "retinanet 16" "yolov3 32" "fcos 4"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/ops_$n.txt)"
done
ODT_TC_MERGE=0 timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/ops_retinanet_16_mg0.txt 2>&1; echo "== retinanet merge=0: $(grep -E 'CUDA-graph' gpurun_out/ops_retinanet_16_mg0.txt)"
grep -n "200x200 7->7\|200x200 28->28\|200x200 16->7" gpurun_out/ops_retinanet_16.txt | head -3 | cut -c1-110
