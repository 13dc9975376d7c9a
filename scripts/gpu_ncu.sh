#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_stem_tc_kernel|nms_per_class" -s 4 -c 2 -o gpurun_out/prof_stem_nms -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu2.log 2>&1; echo "ncu stem/nms exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 56 -c 1 -o gpurun_out/prof_conv1_2 -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu1.log 2>&1; echo "ncu conv exit $?"
