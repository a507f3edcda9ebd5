#!/bin/bash
set -u
mkdir -p gpurun_out /tmp/mb
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart=shared -o /tmp/mb/l2_random4 profiles/microbench/l2_random4.cu > /dev/null 2>&1
timeout 120 /tmp/mb/l2_random4 > gpurun_out/l2_random4.out.txt 2>&1; echo "mb rc=$?"; cat gpurun_out/l2_random4.out.txt
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_pack_sort.py -x -q > gpurun_out/r2_t2.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_t2.log
timeout 1200 python -m pytest tests/test_gpu_scale.py -x -q --durations=8 > gpurun_out/r2_t2_scale.log 2>&1; echo "scale rc=$?"; tail -16 gpurun_out/r2_t2_scale.log
