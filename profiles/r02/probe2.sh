#!/bin/bash
set -u
mkdir -p gpurun_out
RTB_LIB=$PWD/profiles/r02/ab/libr1.so RTB_GB_TRACE=1 timeout 300 python profiles/r02/ab_hash.py > gpurun_out/r2_ab_r1.log 2>&1
RTB_LIB=$PWD/profiles/r02/ab/libr1.so HINT=500000000 timeout 300 python profiles/r02/ab_hash.py > gpurun_out/r2_ab_r1_hint.log 2>&1
RTB_GB_TRACE=1 timeout 300 python profiles/r02/ab_hash.py > gpurun_out/r2_ab_cur.log 2>&1
HINT=500000000 timeout 300 python profiles/r02/ab_hash.py > gpurun_out/r2_ab_cur_hint.log 2>&1
grep -h "^lib" gpurun_out/r2_ab_*.log
grep -h "round 5" gpurun_out/r2_ab_r1.log | tail -2; grep -h "round 5" gpurun_out/r2_ab_cur.log | tail -2
