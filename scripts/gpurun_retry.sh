#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> <script> -- retries while the pod answers "busy" (exit 3; nothing is charged)
log=$1; to=$2; script=$3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "bash $script" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc $rc after $i tries"; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
