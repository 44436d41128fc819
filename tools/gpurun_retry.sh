#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <timeout-seconds> <command...> : retries while the pod answers "transient" (nothing charged)
OUT=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > "$OUT" 2>&1
  if ! grep -q "status=transient" "$OUT"; then exit 0; fi
  sleep 60
done
