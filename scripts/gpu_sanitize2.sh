#!/bin/bash
# compute-sanitizer over the kernels changed late in round 2: NMS rounds behind a named barrier, the staged pointwise
# kernel, the specialised pool kernel, the glue kernels with 32-bit indices.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
CS=/usr/local/cuda/bin/compute-sanitizer
run() {
  local name=$1 tool=$2 flags=$3; shift 3
  timeout 900 $CS --tool $tool $flags --error-exitcode 86 --print-limit 20 \
      python -m pytest "$@" -q -x --timeout 850 -p no:cacheprovider > gpurun_out/r2_sanitize2_$name.log 2>&1
  echo "== $name ($tool $*): exit $? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitize2_$name.log | tail -n 1) ; $(tail -n 1 gpurun_out/r2_sanitize2_$name.log)"
}
run racecheck_tail racecheck "--racecheck-report all" tests/test_gpu_tail.py -k "golden or live_oracle or spill or overflow"
run memcheck_tail memcheck "" tests/test_gpu_tail.py
run memcheck_late memcheck "" tests/test_gpu_conv.py -k "pw or specialised or elementwise or glue or maxpool"
run racecheck_pw racecheck "--racecheck-report all" tests/test_gpu_conv.py -k "conv_pw"
