#!/bin/bash
# usage: tools/runs/retry.sh <timeout> <script> [extra gpurun args]  — retries while the pod answers "busy" (exit 3)
T=$1; S=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" --timeout $T -- "bash $S"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i"; sleep 60
done
exit 3
