#!/bin/bash
# ncu --set full of the two lowest-efficiency launches (conv1_2 = first conv_tc launch, stem)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 70 -c 2 -o gpurun_out/prof_conv1_2 -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu1.log 2>&1; echo "ncu conv exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_stem_tc_kernel -s 2 -c 1 -o gpurun_out/prof_stem -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu2.log 2>&1; echo "ncu stem exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"maxpool|nms_per_class|decode_cand" -s 5 -c 3 -o gpurun_out/prof_glue -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu3.log 2>&1; echo "ncu glue exit $?"
