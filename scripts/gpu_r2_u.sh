#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench exit $?"
python - <<P
import json
j=json.loads(open('gpurun_out/r2u_bench_n1.json').read().strip().splitlines()[-1])
r=j['roofline']; w=j['workloads']['retinanet800_b16']
print('value %.0f (%.3f ms) e2e %.0f | frac_sustained %.3f burst %.3f share %.3f whole %.3f | retina %.0f e2e %.0f frac %.3f whole %.3f share %.3f' % (j['value'], j['ms_per_step'], j['e2e']['value'], r['frac_sustained'], r['frac_burst'], r['share_of_step'], r['whole_step_frac_sustained'], w['value'], w['e2e']['value'], w['roofline']['frac_sustained'], w['roofline']['whole_step_frac_sustained'], w['roofline']['share_of_step']))
print(j['clocks'], j['e2e']['clocks'])
P
