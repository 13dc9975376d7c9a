#!/bin/bash
# round 2, call G: NMS with the short kernel completing images; training step on the GPU
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/r2g_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "autograd loss" gpurun_out/r2g_gpu_tests.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2g_gpu_tests.log | tail -n 15
for short in 0 1 0 1; do
  for m in "ssd300 64" "retinanet 16" "yolov3 32"; do
    echo "NMS_SHORT=$short $(ODT_NMS_SHORT=$short timeout 300 python scripts/tail_micro.py $m 2>&1 | tail -n 1)"
  done
done
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench exit $?"; tail -n 3 gpurun_out/r2g_bench_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e'].get('clocks'),'frac',d['roofline']['frac_sustained'],d['roofline']['whole_step_frac_sustained'],d['clocks'])
print('ssd tail',{k:d['tail_roofline_ssd300'][k] for k in ('decode_us','nms_us','frac_of_hbm','decode_frac_of_hbm','launch_floor_us')})
w=d['workloads']['retinanet800_b16']; print('retina',w['value'],w['ms_per_step'],w['e2e']['value'],w['roofline']['whole_step_frac_sustained'],w['clocks'],w['e2e']['clocks'])
print('retina tail',{k:d['tail_roofline'][k] for k in ('decode_us','nms_us','frac_of_hbm','decode_frac_of_hbm')})
P
