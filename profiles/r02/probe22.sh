#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python profiles/bench_reference_grid.py > gpurun_out/r02_reference_grid.jsonl 2> gpurun_out/r02_reference_grid.err; echo "grid rc=$?"
timeout 600 python profiles/bench_reference_lexsort.py > gpurun_out/r02_reference_lexsort.jsonl 2> gpurun_out/r02_reference_lexsort.err; echo "lexsort rc=$?"
head -12 gpurun_out/r02_reference_grid.jsonl; tail -4 gpurun_out/r02_reference_lexsort.jsonl
