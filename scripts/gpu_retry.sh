#!/bin/bash
# usage: scripts/gpu_retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while the pod answers "busy", exit code 3)
log=$1; shift; to=$1; shift
cd /root/repo
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  exit $rc
done
exit 3
