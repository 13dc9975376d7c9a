#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_candidates -s 2 -c 1 -o gpurun_out/prof_decode_yolo -f python scripts/tail_micro.py yolov3 32 > gpurun_out/ncu4.log 2>&1; echo "ncu exit $?"
