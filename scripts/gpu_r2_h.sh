#!/bin/bash
# per-kernel durations of the tail (ncu launch list) with the warp-per-list NMS kernel on / off + tail tests
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_tail.py tests/test_gpu_models.py -q -m gpu --timeout 600 2>&1 | tail -n 4
for m in "ssd300 64" "yolov3 32" "retinanet 16"; do
for s in 1 0; do
n=$(echo $m | tr ' ' '_')
ODT_NMS_SHORT=$s timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nms|decode" -c 12 --csv --log-file gpurun_out/r2h_tail_${n}_short$s.csv python scripts/tail_micro.py $m > gpurun_out/r2h_micro.log 2>&1
python - <<P
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2h_tail_${n}_short$s.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki][:28]].append(float(r[vi].replace(',','')))
print('$m SHORT=$s', {k:(len(v), round(sum(v)/len(v)/1000,2)) for k,v in d.items()})
P
done
echo "$m graph-timed: $(timeout 300 python scripts/tail_micro.py $m 2>&1 | tail -n 1)"
done
