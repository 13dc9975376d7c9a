#!/bin/bash
# round 2, call C: new decode / NMS kernels, RGBX stems, pool+BN fusion: tests, A/B, per-op tables, bench
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/r2c_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "rows max|classes clean|end-to-end" gpurun_out/r2c_gpu_tests.log | head -40; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c_gpu_tests.log | tail -n 15
for occ in 3 4 3 4; do
  for m in "ssd300 64" "retinanet 16" "yolov3 32"; do
    echo "OCC=$occ $(ODT_DECODE_OCC=$occ timeout 300 python scripts/tail_micro.py $m 2>&1 | tail -n 1)"
  done
done
for m in "ssd300 64" "retinanet 16"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2c_ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/r2c_ops_$n.txt)"; grep -E "^decode" gpurun_out/r2c_ops_$n.txt
  ODT_STEM_RGBX=0 timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2c_ops_${n}_norgbx.txt 2>&1; echo "== $m (fp32-gather stem): $(grep -E 'CUDA-graph' gpurun_out/r2c_ops_${n}_norgbx.txt)"
done
head -n 8 gpurun_out/r2c_ops_retinanet_16.txt
head -n 4 gpurun_out/r2c_ops_ssd300_64.txt
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench exit $?"; tail -n 3 gpurun_out/r2c_bench_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e'].get('clocks'),'frac',d['roofline']['frac_sustained'],d['roofline']['whole_step_frac_sustained'],d['clocks'])
print('ssd tail',{k:d['tail_roofline_ssd300'][k] for k in ('decode_us','nms_us','frac_of_hbm','decode_frac_of_hbm','launch_floor_us')})
w=d['workloads']['retinanet800_b16']; print('retina',w['value'],w['ms_per_step'],w['e2e']['value'],w['roofline']['whole_step_frac_sustained'])
print('retina tail',{k:d['tail_roofline'][k] for k in ('decode_us','nms_us','frac_of_hbm','decode_frac_of_hbm')})
P
for e in rn r12n; do echo "1x1 7->28 extra=$e: $(ODT_MICRO_EXTRA=$e timeout 120 python scripts/conv_micro.py 16 200 200 7 28 1 1 0 0 0 50 | tail -n 1)"; done
ODT_MICRO_EXTRA=r12n timeout 300 ncu --set full --clock-control none --import-source on -s 3 -c 1 -o gpurun_out/r2c_prof_1x1_7_28 -f python scripts/conv_micro.py 16 200 200 7 28 1 1 0 0 0 3 > gpurun_out/r2c_ncu_1x1.log 2>&1; echo "ncu 1x1 exit $?"
