#!/bin/bash
# 2-GPU validation of the sharded path (torchrun, NCCL all-gather of detection records)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; cat gpurun_out/bench_n2.json; tail -n 5 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/check_sharded.py > gpurun_out/sharded.log 2>&1; echo "sharded check exit $?"; tail -n 6 gpurun_out/sharded.log
