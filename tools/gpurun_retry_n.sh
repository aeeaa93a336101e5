#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout> "<command>" <log>   — retries while the pod answers busy, at most 30 times
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3" > "$4" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$4"; then echo "done rc=$rc" >> "$4"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$4"
