#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for w in 1 0; do for a in 0 1; do
  ODT_TC_DEBUG_ALIGNED=$a ODT_TC_WRES=$w python scripts/conv_micro.py 64 300 300 64 64 3 1 2 | sed "s/^/aligned=$a /"
  ODT_TC_DEBUG_ALIGNED=$a ODT_TC_WRES=$w python scripts/conv_micro.py 64 150 150 64 128 3 1 0 | sed "s/^/aligned=$a /"
  ODT_TC_DEBUG_ALIGNED=$a ODT_TC_WRES=$w python scripts/conv_micro.py 64 150 150 128 128 3 1 2 | sed "s/^/aligned=$a /"
done; done
