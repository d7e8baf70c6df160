#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3: nothing charged).  usage: tools/gpu.sh <timeout_s> '<command>'
T="$1"; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu.sh] transient (attempt $attempt), retrying in 60 s"
  sleep 60
done
exit 3
