#!/bin/bash
# glue kernels with 32-bit index math + vectorised parameter loads: parity tests, RetinaNet / YOLOv3 / FCOS graphs
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -q -k "pool or glue or elementwise or upsample or retinanet or yolo or fcos" --timeout 300 > gpurun_out/r2s_tests.log 2>&1
echo "tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2s_tests.log | tail -n 6
for m in "retinanet 16" "yolov3 32" "fcos 4"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2s_ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/r2s_ops_$n.txt) $(grep -E 'Upsample|PoolOp' gpurun_out/r2s_ops_$n.txt | awk '{printf "%s ", $(NF-2)}')"
done
