#!/bin/bash
set -u
mkdir -p gpurun_out
for sp in 0 1; do RTB_GB_SMALL=$sp timeout 300 python profiles/r02/small_latency2.py > gpurun_out/r2_small3_$sp.json 2>&1; cat gpurun_out/r2_small3_$sp.json; done
timeout 1500 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/r2_t6_all.log 2>&1; echo "all gpu tests rc=$?"; tail -24 gpurun_out/r2_t6_all.log
