#!/bin/bash
# ncu --set full capture of one quick_bench case (run under gpurun, one GPU).
#   tools/ncu_capture.sh <out name> <kernel regex> <quick_bench case>
set -e
out=$1; regex=$2; shift 2
mkdir -p gpurun_out
REPS=1 ncu --set full --clock-control none --import-source on -k regex:$regex \
  -s 2 -c 1 -f -o gpurun_out/$out python tools/quick_bench.py "$@" \
  > gpurun_out/$out.log 2>&1 || tail -5 gpurun_out/$out.log
