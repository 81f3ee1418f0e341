#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).
#   scripts/gpurun_retry.sh <log file> <gpurun args...>
log="$1"; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ "$rc" != 3 ]; then exit "$rc"; fi
  sleep 90
done
exit 3
