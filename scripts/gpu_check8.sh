#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in tail conv models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/test_$t.log | tail -n 12
done
for m in "ssd300 64" "retinanet 16" "yolov3 32" "fcos 4" "ssd300 1"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$n.txt 2>&1; echo "== $m"; grep -E "CUDA-graph|tail \(|decode " gpurun_out/ops_$n.txt
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['clocks'],d.get('cpu_baseline'))"; tail -n 3 gpurun_out/bench.err
