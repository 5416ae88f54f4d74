#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gprun_retry.sh <timeout> [--gpus N] <command>
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun --timeout $T $G -- "$@" > /tmp/gprun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gprun_last.txt || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
cat /tmp/gprun_last.txt
