#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_conv.py -q -m gpu --timeout 300 -k "fcos or glue or group" > gpurun_out/test_fcos.log 2>&1
echo "fcos/glue tests exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_fcos.log | tail -n 8
for k in 1 0; do ODT_GN_FUSED=$k timeout 600 python scripts/profile_ops.py fcos 4 > gpurun_out/ops_fcos_4_gn$k.txt 2>&1; echo "== fcos GN_FUSED=$k: $(grep -E 'CUDA-graph' gpurun_out/ops_fcos_4_gn$k.txt)"; done
