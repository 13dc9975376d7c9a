#!/bin/bash
# NMS per-round clock64 timeline of one block (debug build) + warm timings with ODT_NMS_ADAPT=0/1
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for m in "ssd300 1" "ssd300 64" "retinanet 16" "yolov3 32"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/nms_timeline.py $m > gpurun_out/r2n_nms_timeline_$n.txt 2>&1; echo "== $m"; grep -E "candidates|graph of 50|ADAPT|round  [0-3]:|Error" gpurun_out/r2n_nms_timeline_$n.txt
done
