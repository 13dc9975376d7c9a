#!/bin/bash
# 2-GPU validation of the sharded path: torchrun, image shards, ONE all-gather of the packed detection records
# (C-ABI communicator odt_ctx / odt_allgather_dets by default, torch.distributed with ODT_COLLECTIVES=torch)
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_multi.sh'
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -k "sharded_two" -s 2>&1 | tail -n 4
for coll in cabi torch; do
  ODT_COLLECTIVES=$coll timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/check_sharded.py > gpurun_out/r2_sharded_$coll.log 2>&1; echo "sharded check ($coll) exit $?"; tail -n 3 gpurun_out/r2_sharded_$coll.log
done
for coll in cabi torch; do
  ODT_COLLECTIVES=$coll timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_$coll.json 2> gpurun_out/r2_bench_n2_$coll.err; echo "bench n2 ($coll) exit $?"; tail -n 3 gpurun_out/r2_bench_n2_$coll.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n2_$coll.json').read().strip().splitlines()[-1])
print('$coll: value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'d2h',d['e2e']['d2h_bytes_per_step'],'retina',d['workloads']['retinanet800_b16']['value'],d['workloads']['retinanet800_b16']['e2e']['value'])
P
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n1_samebox.json 2>/dev/null; python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n1_samebox.json').read().strip().splitlines()[-1])
print('n1 same box: value',d['value'],'e2e',d['e2e']['value'],'retina',d['workloads']['retinanet800_b16']['value'],d['workloads']['retinanet800_b16']['e2e']['value'])
P
