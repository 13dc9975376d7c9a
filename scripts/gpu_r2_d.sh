#!/bin/bash
# round 2, call D: halo'd extra outputs, fixed decode, tests, per-op tables, decode ncu with source, bench
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/r2d_gpu_tests.log 2>&1
echo "pytest -m gpu exit $?"; grep -E "rows max|classes clean|end-to-end" gpurun_out/r2d_gpu_tests.log | cut -c1-400 | head -40; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2d_gpu_tests.log | tail -n 15
for m in "retinanet 16" "ssd300 64" "yolov3 32" "fcos 4"; do
  n=$(echo $m | tr ' ' '_')
  timeout 600 python scripts/profile_ops.py $m > gpurun_out/r2d_ops_$n.txt 2>&1; echo "== $m: $(grep -E 'CUDA-graph' gpurun_out/r2d_ops_$n.txt)"; grep -E "^decode" gpurun_out/r2d_ops_$n.txt
done
ODT_HALO_AUX=0 timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/r2d_ops_retinanet_16_noaux.txt 2>&1; echo "== retinanet (ODT_HALO_AUX=0): $(grep -E 'CUDA-graph' gpurun_out/r2d_ops_retinanet_16_noaux.txt)"
head -n 40 gpurun_out/r2d_ops_retinanet_16.txt | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_candidates_kernel -s 2 -c 1 -o gpurun_out/r2d_prof_decode_retina -f python scripts/tail_micro.py retinanet 16 > gpurun_out/r2d_ncu_decode.log 2>&1; echo "ncu decode exit $?"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench exit $?"; tail -n 3 gpurun_out/r2d_bench_n1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2d_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e'].get('clocks'),'frac',d['roofline']['frac_sustained'],d['roofline']['whole_step_frac_sustained'],d['clocks'])
print('ssd tail',{k:d['tail_roofline_ssd300'][k] for k in ('decode_us','nms_us','frac_of_hbm','decode_frac_of_hbm','launch_floor_us')})
w=d['workloads']['retinanet800_b16']; print('retina',w['value'],w['ms_per_step'],w['e2e']['value'],w['roofline']['whole_step_frac_sustained'])
print('retina tail',{k:d['tail_roofline'][k] for k in ('decode_us','nms_us','frac_of_hbm','decode_frac_of_hbm')})
P
