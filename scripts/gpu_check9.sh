#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in conv tail models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error|assert" gpurun_out/test_$t.log | tail -n 14
done
timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/ops_ssd300_64.txt 2>&1; echo "== ssd300 64 halo"; grep -E "conv1_|conv2_|pool|CUDA-graph" gpurun_out/ops_ssd300_64.txt | head -12
ODT_HALO=0 timeout 600 python scripts/profile_ops.py ssd300 64 2>&1 | grep -E "CUDA-graph"
timeout 600 python scripts/profile_ops.py yolov3 32 > gpurun_out/ops_yolov3_32.txt 2>&1; echo "== yolo halo"; grep -E "CUDA-graph" gpurun_out/ops_yolov3_32.txt; head -12 gpurun_out/ops_yolov3_32.txt | cut -c1-110
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['clocks'])"; tail -n 3 gpurun_out/bench.err
