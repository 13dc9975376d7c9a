#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for m in "ssd300 64" "retinanet 16" "yolov3 32" "fcos 4"; do
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$(echo $m | cut -d' ' -f1).txt 2>&1; echo "== $m"; grep -E "CUDA-graph|tail \(|decode " gpurun_out/ops_$(echo $m | cut -d' ' -f1).txt
done
