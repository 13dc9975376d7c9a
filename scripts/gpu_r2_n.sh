#!/bin/bash
# NMS: adaptive number of participating warps (A/B) + per-round clock64 timeline of one block
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_tail.py -q -x --timeout 300 > gpurun_out/r2n_tail_tests.log 2>&1
echo "tail tests exit $?"; tail -n 3 gpurun_out/r2n_tail_tests.log
for m in "ssd300 1" "ssd300 64" "retinanet 16" "yolov3 32"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/nms_timeline.py $m > gpurun_out/r2n_nms_timeline_$n.txt 2>&1; echo "== $m"; grep -E "candidates|graph of 50|ADAPT" gpurun_out/r2n_nms_timeline_$n.txt
done
for ad in 0 1; do
  ODT_NMS_ADAPT=$ad timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/r2n_ops_ssd300_64_adapt$ad.txt 2>&1; echo "== ssd300 64 adapt=$ad: $(grep -E 'CUDA-graph|^decode' gpurun_out/r2n_ops_ssd300_64_adapt$ad.txt | tr '\n' ' ')"
  ODT_NMS_ADAPT=$ad timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/r2n_ops_retinanet_16_adapt$ad.txt 2>&1; echo "== retinanet 16 adapt=$ad: $(grep -E 'CUDA-graph|^decode' gpurun_out/r2n_ops_retinanet_16_adapt$ad.txt | tr '\n' ' ')"
done
