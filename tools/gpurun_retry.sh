#!/bin/bash
# gpurun with retries while the pod answers busy / transient (exit 3, or "transient" in the verdict): tools/gpurun_retry.sh <log> [gpurun args...]
log=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc = 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
