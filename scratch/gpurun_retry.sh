#!/usr/bin/env bash
# usage: scratch/gpurun_retry.sh <logfile> [gpurun args...] -- '<command>'
# retries while the pod answers "busy / draining" (nothing charged), up to ~40 minutes
log="$1"; shift
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 170
done
exit 3
