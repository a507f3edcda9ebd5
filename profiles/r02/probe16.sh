#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_groupby.py -x -q -k "adopt_seed or stream" > gpurun_out/r02_t_adopt.log 2>&1; echo "adopt rc=$?"; tail -5 gpurun_out/r02_t_adopt.log
bash profiles/r02/probe14.sh 2
