#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -c 140 --csv --log-file gpurun_out/conv_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1; echo "ncu traffic exit $?"
python scripts/conv_traffic.py gpurun_out/conv_traffic.csv 28 gpurun_out/conv_traffic.json
