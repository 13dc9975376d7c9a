#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 120 -x > gpurun_out/test_conv.log 2>&1
echo "conv tests exit $?"; grep -E "passed|failed|FAILED|Error|assert|rror" gpurun_out/test_conv.log | tail -n 20
for pr in 1 0; do
  ODT_TC_FLAT_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 300 300 64 64 3 1 2 | sed "s/^/FPAIR=$pr /"
  ODT_TC_FLAT_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 150 150 64 128 3 1 0 | sed "s/^/FPAIR=$pr /"
  ODT_TC_FLAT_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 150 150 128 128 3 1 2 | sed "s/^/FPAIR=$pr /"
  ODT_TC_FLAT_PAIR=$pr timeout 120 python scripts/conv_micro.py 16 200 200 28 28 3 1 0 | sed "s/^/FPAIR=$pr /"
  ODT_TC_FLAT_PAIR=$pr timeout 120 python scripts/conv_micro.py 32 208 208 32 64 3 1 0 | sed "s/^/FPAIR=$pr /"
done
ODT_TC_WRES=0 timeout 120 python scripts/conv_micro.py 64 150 150 128 128 3 1 2 | sed "s/^/FPAIR=1 /"
ODT_TC_WRES=0 timeout 120 python scripts/conv_micro.py 64 150 150 64 128 3 1 0 | sed "s/^/FPAIR=1 /"
