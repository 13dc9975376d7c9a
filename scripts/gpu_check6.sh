#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for t in tail conv models; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --timeout 300 > gpurun_out/test_$t.log 2>&1
  echo "test_gpu_$t exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/test_$t.log | tail -n 12
done
timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/ops_retinanet.txt 2>&1; grep -E "tail|decode|sum of" gpurun_out/ops_retinanet.txt
timeout 600 python scripts/profile_ops.py yolov3 32 > gpurun_out/ops_yolov3.txt 2>&1; grep -E "tail|decode|sum of|Nearest" gpurun_out/ops_yolov3.txt
timeout 600 python scripts/profile_ops.py fcos 4 > gpurun_out/ops_fcos.txt 2>&1; python - <<'PY'
import re,collections
agg=collections.defaultdict(float)
for ln in open('gpurun_out/ops_fcos.txt'):
    m=re.match(r"\d+\s+(\S+)\s+.*?\s+([\d.]+)\s+[\d.]+\s+[\d.]+$", ln.rstrip())
    if m: agg[m.group(1)]+=float(m.group(2))
print(dict(agg))
PY
tail -n 2 gpurun_out/ops_fcos.txt
