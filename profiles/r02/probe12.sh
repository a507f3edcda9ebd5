#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python profiles/r02/sanitize_smoke.py > gpurun_out/r02_san_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r02_san_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/r02/sanitize_smoke.py > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r02_memcheck.log
