#!/bin/bash
# two-stage pipeline (tail of step i under the body of step i+1): parity, then bench A/B on one box
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_models.py -q -k "stream or pipelined or sharded" --timeout 300 > gpurun_out/r2t_tests.log 2>&1
echo "tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/r2t_tests.log | tail -n 8
for rep in 1 2; do for pl in 0 1; do
  ODT_PIPELINE=$pl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_pl${pl}_$rep.json 2> gpurun_out/r2t_bench_pl${pl}_$rep.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2t_bench_pl${pl}_$rep.json').read().strip().splitlines()[-1])
    w=d['workloads']['retinanet800_b16']
    print('pipeline=$pl rep $rep: ssd300 value %.0f (%.3f ms) e2e %.0f clocks %s | retinanet value %.0f (%.3f ms) e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], w['value'], w['ms_per_step'], w['e2e']['value']))
except Exception as e:
    print('pipeline=$pl rep $rep: FAILED', e); print(open('gpurun_out/r2t_bench_pl${pl}_$rep.err').read()[-1500:])
P
done; done
