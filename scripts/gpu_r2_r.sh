#!/bin/bash
# specialised fp16 max-pool kernel: parity + same-box A/B on the SSD graphs; ncu --set full of a RetinaNet P3 tower layer
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_conv.py -q -k "pool or glue or elementwise" --timeout 300 > gpurun_out/r2r_pool_tests.log 2>&1
echo "pool tests exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2r_pool_tests.log | tail -n 6
for rep in 1 2; do for f in 0 1; do
  ODT_POOL_FAST=$f timeout 600 python scripts/profile_ops.py ssd300 64 > gpurun_out/r2r_ops_ssd300_64_pool${f}_$rep.txt 2>&1
  echo "== ssd300 64 pool_fast=$f rep $rep: $(grep -E 'CUDA-graph' gpurun_out/r2r_ops_ssd300_64_pool${f}_$rep.txt) $(grep PoolOp gpurun_out/r2r_ops_ssd300_64_pool${f}_$rep.txt | awk '{printf "%s ", $(NF-2)}')"
done; done
for f in 0 1; do
  ODT_POOL_FAST=$f timeout 600 python scripts/profile_ops.py ssd512 32 > gpurun_out/r2r_ops_ssd512_32_pool$f.txt 2>&1
  echo "== ssd512 32 pool_fast=$f: $(grep -E 'CUDA-graph' gpurun_out/r2r_ops_ssd512_32_pool$f.txt)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2r_tower_p3 -f python scripts/conv_micro.py 16 100 100 256 256 3 1 0 1 1 3 > gpurun_out/r2r_ncu_tower.log 2>&1; echo "ncu tower exit $?"
