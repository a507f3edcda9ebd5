#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl_abi.py -x -q > gpurun_out/r02_t_nccl.log 2>&1; echo "nccl abi rc=$?"; tail -15 gpurun_out/r02_t_nccl.log
