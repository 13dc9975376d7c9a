#!/bin/bash
# First GPU call of round 2: settle the two opt-in kernels of round 1 on ONE box (same clocks, same process
# environment), then the timeline of a thin layer.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_gpu_tests_default.log 2>&1
echo "pytest -m gpu (defaults) exit $?"; tail -n 3 gpurun_out/r2_gpu_tests_default.log
ODT_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_conv.py -q -k "tapn or thin" > gpurun_out/r2_tapn_tests.log 2>&1
echo "tapn/thin tests exit $?"; tail -n 2 gpurun_out/r2_tapn_tests.log
{
  # conv1_2 of SSD300 (pooled), conv1_2 of SSD512, a thin RetinaNet layer (7->7 and 28->28 at 200x200, B=16),
  # YOLOv3's 32->64 and FCOS's 64->64; each line: default path, then taps-as-N, three interleaved repeats
  for shape in "64 300 300 64 64 3 1 2 1 1" "32 512 512 64 64 3 1 2 1 1" "16 200 200 7 7 3 1 0 1 1" \
               "16 200 200 28 28 3 1 0 1 1" "32 208 208 32 64 3 1 0 1 1" "4 256 256 64 64 3 1 0 1 1" \
               "32 416 416 32 64 3 2 0 0 0" "16 200 200 16 7 1 1 0 0 0" "16 200 200 7 28 1 1 0 0 0" \
               "16 100 100 14 14 3 1 0 1 1"; do   # the last one: YOLOv3 block1 (stride 2, im2col mode): K-skip only
    for rep in 1 2 3; do
      ODT_TC_TAPN=0 ODT_TC_KSKIP=0 python scripts/conv_micro.py $shape 50 | sed 's/^/base   /'
      ODT_TC_TAPN=0 ODT_TC_KSKIP=1 python scripts/conv_micro.py $shape 50 | sed 's/^/kskip  /'
      ODT_TC_TAPN=2 python scripts/conv_micro.py $shape 50 | sed 's/^/tapn   /'
      ODT_TC_THIN=2 python scripts/conv_micro.py $shape 50 | sed 's/^/thin   /'   # = base unless Cin, Cout <= 16 (3x3) / 32 (1x1)
    done
  done
} > gpurun_out/r2_ab_micro.txt 2>&1
tail -n 60 gpurun_out/r2_ab_micro.txt
ODT_TC_TAPN=1 ODT_TC_KSKIP=1 ODT_TC_THIN=1 timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2_gpu_tests_tapn.log 2>&1
echo "pytest -m gpu with TAPN=KSKIP=1 exit $?"; tail -n 3 gpurun_out/r2_gpu_tests_tapn.log
for t in 0 1; do
  for m in "ssd300 64" "retinanet 16"; do
    n=$(echo $m | tr ' ' '_')
    ODT_TC_TAPN=$t ODT_TC_KSKIP=$t ODT_TC_THIN=$t timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2_ops_${n}_tapn$t.txt 2>&1
    echo "== $m TAPN=KSKIP=$t: $(grep -E 'CUDA-graph' gpurun_out/r2_ops_${n}_tapn$t.txt)"
  done
done
for t in 0 1 0 1; do
  ODT_TC_TAPN=$t ODT_TC_KSKIP=$t timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_tapn$t.json 2> gpurun_out/r2_bench_tapn$t.err
  echo "bench TAPN=KSKIP=$t: $(python -c "import json,sys; d=json.loads(open('gpurun_out/r2_bench_tapn$t.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])")"
done
timeout 300 python scripts/tc_timeline.py 16 200 200 7 7 3 1 0 1 1 > gpurun_out/r2_timeline_thin.txt 2>&1; echo "timeline exit $?"
ODT_TC_KSKIP=1 timeout 300 python scripts/tc_timeline.py 16 200 200 7 7 3 1 0 1 1 > gpurun_out/r2_timeline_thin_kskip.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tapn_kernel -s 3 -c 1 -o gpurun_out/r2_prof_tapn_c12 -f env ODT_TC_TAPN=1 python scripts/conv_micro.py 64 300 300 64 64 3 1 2 1 1 3 > gpurun_out/r2_ncu_tapn.log 2>&1; echo "ncu tapn exit $?"
ODT_BENCH_DEFERRED=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_deferred.json 2> gpurun_out/r2_bench_deferred.err; echo "bench deferred e2e: $(python -c "import json; d=json.loads(open(\"gpurun_out/r2_bench_deferred.json\").read().strip().splitlines()[-1]); print(d[\"value\"], d[\"e2e\"])")"
ODT_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_models.py -q -k deferred 2>&1 | tail -n 2
ODT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -k full_size 2>&1 | tail -n 3
