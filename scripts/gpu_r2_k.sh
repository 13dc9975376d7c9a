#!/bin/bash
# round-2 closing run: full GPU suite, smoke, bench line, per-op tables, ncu launch list + conv DRAM traffic, conv_pw A/B + ncu
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/r2k_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2k_gpu_tests.log | tail -n 8
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; tail -c 2500 gpurun_out/r2k_bench_n1.json
for rep in 1 2; do for pw in 0 1 2; do
  ODT_TC_PW=$pw timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/r2k_ops_retinanet_16_pw${pw}_$rep.txt 2>&1
  echo "== retinanet 16 pw=$pw rep $rep: $(grep -E 'CUDA-graph' gpurun_out/r2k_ops_retinanet_16_pw${pw}_$rep.txt)"
done; done
for m in "ssd300 64" "ssd512 32" "yolov3 32" "fcos 4" "ssd300 1"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2k_ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/r2k_ops_$n.txt)"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_|decode_|nms_|pool|l2norm|pack_|affine|upsample|normalize" -c 600 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-retinanet > gpurun_out/r2k_bench_under_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:conv_(tc|tapn|thin|pw)_kernel" -c 140 --csv --log-file gpurun_out/r2k_conv_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-retinanet > gpurun_out/r2k_bench_under_ncu2.log 2>&1; echo "ncu traffic exit $?"
python scripts/conv_traffic.py gpurun_out/r2k_conv_traffic.csv 28 gpurun_out/r2k_conv_traffic.json
ODT_TC_PW=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pw_kernel -c 4 -o gpurun_out/r2k_pw -f python scripts/profile_ops.py retinanet 16 > gpurun_out/r2k_ncu_pw.log 2>&1; echo "ncu pw exit $?"
