#!/bin/bash
# staged pointwise kernel (conv_pw.cu): parity through the C ABI, then same-box A/B on the whole RetinaNet / FCOS graphs
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_gpu_conv.py -q -x -k "pw or thin" --timeout 300 > gpurun_out/r2l_pw_tests.log 2>&1
echo "pw tests exit $?"; tail -n 4 gpurun_out/r2l_pw_tests.log
for rep in 1 2; do
for pw in 0 1; do
  ODT_TC_PW=$pw timeout 600 python scripts/profile_ops.py retinanet 16 > gpurun_out/r2l_ops_retinanet_16_pw${pw}_$rep.txt 2>&1
  echo "== retinanet 16 pw=$pw rep $rep: $(grep -E 'CUDA-graph' gpurun_out/r2l_ops_retinanet_16_pw${pw}_$rep.txt)"
done
done
for pw in 0 1; do
  ODT_TC_PW=$pw timeout 600 python scripts/profile_ops.py fcos 4 > gpurun_out/r2l_ops_fcos_4_pw$pw.txt 2>&1
  echo "== fcos 4 pw=$pw: $(grep -E 'CUDA-graph' gpurun_out/r2l_ops_fcos_4_pw$pw.txt)"
done
timeout 600 python -m pytest tests/test_gpu_models.py -q -x -k "retinanet" --timeout 300 > gpurun_out/r2l_retina_tests.log 2>&1
echo "retinanet model tests exit $?"; tail -n 3 gpurun_out/r2l_retina_tests.log
