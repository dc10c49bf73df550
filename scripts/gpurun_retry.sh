#!/bin/bash
# scripts/gpurun_retry.sh <timeout> [--gpus N] -- <command>: gpurun, retried while the pod answers "transient"/busy
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=ok\|status=fail\|status=timeout\|status=error"; then exit 0; fi
  if ! echo "$out" | grep -q "status=transient\|rc=3\|busy\|retry in a few minutes"; then exit 0; fi
  echo "[retry $attempt] waiting 200 s"; sleep 200
done
