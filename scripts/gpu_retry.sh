#!/bin/bash
# usage: scripts/gpu_retry.sh LOGFILE [gpurun args...] -- '<command>'   retries while the pod answers busy / transient (exit 3)
LOG=$1; shift
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" >/dev/null
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$LOG" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "gpurun finished rc=$rc" >> "$LOG"
