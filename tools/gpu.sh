#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged):  tools/gpu.sh LOGFILE TIMEOUT [--gpus N] 'command'
LOG=$1; TMO=$2; shift 2
EXTRA=""
if [ "$1" = "--gpus" ]; then EXTRA="--gpus $2"; shift 2; fi
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $EXTRA --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 75
done
exit 3
