#!/bin/bash
# N-GPU validation of the sharded path after the side-stream all-gather change:  gpurun --gpus N -- 'bash scripts/gpu_multi_r2.sh N'
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/gpus_n$N.txt
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -k "sharded" -s 2>&1 | tail -n 4
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/check_sharded.py > gpurun_out/r2_sharded_n2.log 2>&1; echo "sharded check exit $?"; tail -n 3 gpurun_out/r2_sharded_n2.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N exit $?"; tail -n 3 gpurun_out/r2_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('n$N: value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'d2h',d['e2e']['d2h_bytes_per_step'],'retina',d['workloads']['retinanet800_b16']['value'],d['workloads']['retinanet800_b16']['e2e']['value'])
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; echo "reference arm n$N exit $?"; cut -c1-400 gpurun_out/r2_bench_ref_n$N.json
