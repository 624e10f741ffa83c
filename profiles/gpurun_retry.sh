#!/usr/bin/env bash
# usage: [GPUS=2] profiles/gpurun_retry.sh <timeout_s> <command...>  — retries while the pod answers "busy" (nothing is charged for those)
T="$1"; shift
EXTRA=()
if [[ -n "${GPUS:-}" ]]; then EXTRA=(--gpus "$GPUS"); fi
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "${EXTRA[@]}" -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box\|busy"; then sleep 150; continue; fi
  break
done
