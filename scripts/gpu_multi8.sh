#!/bin/bash
# 8-GPU validation of the sharded bench path (torchrun, NCCL all-gather of detection records)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/gpus8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 exit $?"; cat gpurun_out/bench_n8.json | cut -c1-700; tail -n 3 gpurun_out/bench_n8.err | cut -c1-300
