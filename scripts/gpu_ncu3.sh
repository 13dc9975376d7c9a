#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_stem_tc_kernel" -s 4 -c 1 -o gpurun_out/prof_stem2 -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu3.log 2>&1; echo "ncu stem exit $?"
