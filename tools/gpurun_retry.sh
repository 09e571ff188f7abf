#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit code 3 / status=transient); usage: gpurun_retry.sh TIMEOUT 'command'
for try in $(seq 1 40); do
  out=$(gpurun --timeout $1 -- "$2" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -${TAILN:-25}; exit $rc
done
echo "gpurun_retry: still busy after 40 tries"; exit 3
