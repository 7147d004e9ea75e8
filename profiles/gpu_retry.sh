#!/bin/bash
# profiles/gpu_retry.sh TIMEOUT SCRIPT [LOG] -- run `bash SCRIPT` under gpurun, retrying while the pod has no free slot
# (a refused attempt costs nothing).  Tooling for this repo's own GPU sessions; not part of the product.
t=$1; s=$2; log=${3:-gpurun_out/$(basename $s .sh)_call.log}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_ARGS --timeout $t -- "bash $s" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
