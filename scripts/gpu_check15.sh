#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 300 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 120 -x -k "pairs" > gpurun_out/test_pairs.log 2>&1
echo "pair tests exit $?"; grep -E "passed|failed|FAILED|Error|assert|rror" gpurun_out/test_pairs.log | tail -n 20
for pr in 1 0; do
  ODT_TC_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 38 38 512 512 3 1 0 0 0 | sed "s/^/PAIR=$pr /"
  ODT_TC_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 75 75 256 256 3 1 0 0 0 | sed "s/^/PAIR=$pr /"
  ODT_TC_PAIR=$pr timeout 120 python scripts/conv_micro.py 16 100 100 256 256 3 1 0 0 0 | sed "s/^/PAIR=$pr /"
  ODT_TC_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 19 19 1024 1024 1 1 0 0 0 | sed "s/^/PAIR=$pr /"
  ODT_TC_PAIR=$pr timeout 120 python scripts/conv_micro.py 64 75 75 128 256 3 1 0 0 0 | sed "s/^/PAIR=$pr /"
done
