#!/bin/bash
# One gpurun call: GPU parity tests (each file in its own process so a trapped
# kernel cannot poison the others), smoke, a short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in tail conv models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 -x -s > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; tail -n 25 gpurun_out/test_$t.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 8 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
