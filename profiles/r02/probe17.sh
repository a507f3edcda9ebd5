#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pack_sort.py -x -q -k "lexsort or group_from or glue or cutoffs" > gpurun_out/r02_t_gfl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_t_gfl.log
timeout 900 python profiles/bench_extra.py > gpurun_out/r02_bench_extra2.jsonl 2> gpurun_out/r02_bench_extra2.err; grep "C4\|lexsort" gpurun_out/r02_bench_extra2.jsonl | cut -c1-150
