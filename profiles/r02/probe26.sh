#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r02_t_all3.log 2>&1; echo "all gpu tests rc=$?"; tail -14 gpurun_out/r02_t_all3.log
