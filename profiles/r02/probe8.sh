#!/bin/bash
set -u
mkdir -p gpurun_out
for sp in 0 1; do RTB_GB_SMALL=$sp timeout 300 python profiles/r02/small_latency2.py > gpurun_out/r2_small2_$sp.json 2>&1; cat gpurun_out/r2_small2_$sp.json; done
