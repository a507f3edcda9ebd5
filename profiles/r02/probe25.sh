#!/bin/bash
set -u
mkdir -p gpurun_out
for load in 0.25 0.5 0.7; do
  RTB_GB_LOAD=$load timeout 300 python profiles/trace_c5.py > /dev/null 2> gpurun_out/r02_c5_load$load.log
  echo "== load $load"; sed -n '/second call/,$p' gpurun_out/r02_c5_load$load.log | cut -c1-150
done
