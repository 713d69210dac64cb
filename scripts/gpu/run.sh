#!/bin/bash
# usage: scripts/gpu/run.sh <call-script> [gpurun timeout s] [gpus]
# Runs one scripts/gpu/*.sh call on a B200 box through gpurun, retrying while the pod answers "busy" (exit 3,
# nothing charged).  Output of gpurun goes to /tmp/<name>.out; files the call writes under gpurun_out/ come back.
call=$1; to=${2:-1500}; gpus=${3:-1}
name=$(basename "$call" .sh)
extra=""
[ "$gpus" != "1" ] && extra="--gpus $gpus"
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" $extra -- "bash $call" > /tmp/$name.out 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/$name.out || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
exit $rc
