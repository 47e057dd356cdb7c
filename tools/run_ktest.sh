#!/bin/bash
# Runs every ktest case in its own process (a trapped kernel must not poison the others).
# usage: tools/run_ktest.sh [logfile]   (env NK_GEMM_CTA_GROUP=1 forces single-CTA tiles)
mkdir -p gpurun_out
LOG=${1:-gpurun_out/ktest.log}
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for c in $(./tools/ktest list); do
  timeout 90 ./tools/ktest $c >> $LOG 2>&1
  rc=$?
  if [ $rc -ne 0 ] && [ $rc -ne 1 ]; then echo "$c: EXIT $rc" >> $LOG; fi
done
cat $LOG
