#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in tail conv models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/test_$t.log | tail -n 12
done
timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/ops_ssd300.txt 2>&1; grep -E "conv1_|conv2_|pred|tail|decode|sum of" gpurun_out/ops_ssd300.txt
timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/ops_retinanet.txt 2>&1; grep -E "tail|decode|sum of|7->28|256->189|256->36 " gpurun_out/ops_retinanet.txt | head -20
timeout 600 python scripts/profile_ops.py yolov3 32 2>&1 | grep -E "tail|decode|sum of"
timeout 600 python scripts/profile_ops.py fcos 4 > gpurun_out/ops_fcos.txt 2>&1; python - <<'PY'
import re,collections
agg=collections.defaultdict(float)
for ln in open('gpurun_out/ops_fcos.txt'):
    m=re.match(r"\d+\s+(\S+)\s+.*?\s+([\d.]+)\s+[\d.]+\s+[\d.]+$", ln.rstrip())
    if m: agg[m.group(1)]+=float(m.group(2))
print(dict(agg))
PY
tail -n 2 gpurun_out/ops_fcos.txt
ODT_PDL=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench.err; echo "bench pdl1 exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_pdl1.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"; tail -n 3 gpurun_out/bench.err
ODT_PDL=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench.err; echo "bench pdl0 exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_pdl0.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
