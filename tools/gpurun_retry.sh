#!/usr/bin/env bash
# gpurun with retries while the pod's GPU slots are busy (exit code 3 / "transient": nothing is charged).
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'
t="$1"; g="$2"; shift 2
for attempt in $(seq 1 40); do
    out=$(/usr/local/graft/bin/gpurun --timeout "$t" --gpus "$g" -- "$@" 2>&1)
    rc=$?
    echo "$out" | tail -25
    if echo "$out" | grep -q "status=transient"; then
        echo "[retry] attempt $attempt: slots busy, sleeping 120 s"; sleep 120; continue
    fi
    exit $rc
done
exit 3
