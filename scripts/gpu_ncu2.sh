#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for w in 1 0; do
  ODT_TC_WRES=$w python scripts/conv_micro.py 64 300 300 64 64 3 1 2
  ODT_TC_WRES=$w python scripts/conv_micro.py 64 300 300 64 64 3 1 0
  ODT_TC_WRES=$w python scripts/conv_micro.py 64 150 150 64 128 3 1 0
  ODT_TC_WRES=$w python scripts/conv_micro.py 16 200 200 7 7 3 1 0
  ODT_TC_WRES=$w timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/prof_c12_wres$w -f python scripts/conv_micro.py 64 300 300 64 64 3 1 2 1 1 3 > gpurun_out/ncu_c12_$w.log 2>&1; echo "ncu exit $?"
done
ODT_TC_WRES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/prof_thin -f python scripts/conv_micro.py 16 200 200 7 7 3 1 0 1 1 3 > gpurun_out/ncu_thin.log 2>&1; echo "ncu exit $?"
