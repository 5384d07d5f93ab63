#!/bin/bash
# gpurun with patience: exit code 3 means "no box or slot free right now (nothing charged)" -- wait and ask again.
# usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: pod busy, sleeping 90 s" >&2
  sleep 90
done
exit 3
