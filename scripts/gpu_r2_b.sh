#!/bin/bash
# round 2, call B: new host path + tests + two-workload bench + ncu of the NMS kernel (dense RetinaNet rows)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x -s > gpurun_out/r2b_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "rows max|classes clean|end-to-end" gpurun_out/r2b_gpu_tests.log | head -40; tail -n 5 gpurun_out/r2b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench exit $?"; tail -n 5 gpurun_out/r2b_bench_n1.err; cat gpurun_out/r2b_bench_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_per_class_kernel -s 2 -c 1 -o gpurun_out/r2b_prof_nms_retina -f python scripts/tail_micro.py retinanet 16 > gpurun_out/r2b_ncu_nms.log 2>&1; echo "ncu nms exit $?"; tail -n 4 gpurun_out/r2b_ncu_nms.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_candidates_kernel -s 2 -c 1 -o gpurun_out/r2b_prof_decode_retina -f python scripts/tail_micro.py retinanet 16 > gpurun_out/r2b_ncu_decode.log 2>&1; echo "ncu decode exit $?"
