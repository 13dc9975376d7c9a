#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 56 -c 3 -o gpurun_out/prof_conv1_2 -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu1.log 2>&1; echo "ncu conv exit $?"
