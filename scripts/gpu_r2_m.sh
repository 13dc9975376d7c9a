#!/bin/bash
# conv_pw.cu under ncu: per-launch durations of one RetinaNet forward with the kernel off / on, then a full capture
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for pw in 0 1; do
  ODT_TC_PW=$pw timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/r2m_launches_pw$pw.csv python scripts/profile_ops.py retinanet 16 > gpurun_out/r2m_ncu_pw$pw.log 2>&1; echo "ncu durations pw=$pw exit $?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pw_kernel -c 10 -o gpurun_out/r2m_pw -f python scripts/profile_ops.py retinanet 16 > gpurun_out/r2m_ncu_full.log 2>&1; echo "ncu full exit $?"
