#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_merge.py -x -q > gpurun_out/r2_t3_merge.log 2>&1; echo "merge rc=$?"; tail -25 gpurun_out/r2_t3_merge.log
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_pack_sort.py -q > gpurun_out/r2_t3.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2_t3.log
