#!/bin/bash
# full round validation: GPU tests, per-op tables, bench line, ncu launch list + full captures
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/test_gpu_all.log 2>&1
echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/test_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for m in "ssd300 64" "retinanet 16" "yolov3 32" "fcos 4" "ssd300 1"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/ops_$n.txt)"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1700 gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/prof_pair_c42 -f python scripts/conv_micro.py 64 38 38 512 512 3 1 0 0 0 3 > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair exit $?"
