#!/bin/bash
# usage: gpuretry.sh <timeout> '<command>'   -- retries while gpurun answers "no slot free" (exit 3)
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun ${GPURUN_EXTRA} --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
