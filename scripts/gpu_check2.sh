#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in conv models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 -s > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|max\|err\||loss gpu|end-to-end" gpurun_out/test_$t.log | tail -n 40
done
timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/ops_ssd300.txt 2>&1; echo "profile exit $?"; cat gpurun_out/ops_ssd300.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_ssd300.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 6 -o gpurun_out/prof_conv_tc -f python scripts/profile_ops.py ssd300 64 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
