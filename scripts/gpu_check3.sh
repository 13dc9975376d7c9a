#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in conv tail models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 -s > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/test_$t.log | tail -n 12
done
timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/ops_ssd300.txt 2>&1; echo "profile exit $?"; cat gpurun_out/ops_ssd300.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
