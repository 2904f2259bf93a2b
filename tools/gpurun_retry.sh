#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged).  usage: tools/gpurun_retry.sh LOG [gpurun args...]
log="$1"; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 120
done
