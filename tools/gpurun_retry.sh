#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <script> <log>   — retries while the pod answers busy (exit 3), at most 40 times
for i in $(seq 1 40); do
  gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$3"; then echo "done rc=$rc" >> "$3"; exit $rc; fi
  sleep 120
done
echo "gave up" >> "$3"
