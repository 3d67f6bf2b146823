#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> <command...>   -- retries while the pod answers busy (exit 3)
T=$1; G=$2; shift 2
for n in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
