#!/bin/bash
# round-2 closing validation: full GPU suite, smoke, bench line (N=1), per-op tables of the two bench workloads
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2z_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2z_gpu_tests.log | tail -n 8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 600 gpurun_out/r2z_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "reference arm exit $?"; cut -c1-200 gpurun_out/r2z_bench_ref.json
for m in "ssd300 64" "retinanet 16" "ssd512 32" "yolov3 32" "fcos 4"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2z_ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/r2z_ops_$n.txt)"
done
