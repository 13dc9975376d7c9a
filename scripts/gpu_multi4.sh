#!/bin/bash
# 4-GPU validation of the sharded bench path (torchrun, NCCL all-gather of detection records)
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench n4 exit $?"; cut -c1-600 gpurun_out/bench_n4.json; tail -n 3 gpurun_out/bench_n4.err | cut -c1-300
