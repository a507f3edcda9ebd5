#!/bin/bash
set -u
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err; python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(f"gpurun_out/r2_bench_{sys.argv[1]}.json") if l.startswith("{")][0])
    print(sys.argv[1], "value %.1f G rows/s ms/step %.3f two_call %.3f hash %.3f sum %.3f" % (d["value"]/1e9, d["ms_per_step"], d["two_call"]["ms_per_step"], d["two_call"]["hash_ms_per_step"], d["two_call"]["sum_ms_per_step"]))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
}
run base
RTB_NVCC_FLAGS="-DRTB_STREAM_CS=1" python -m riptable_b200.build --force > /dev/null 2>&1; run cs
RTB_NVCC_FLAGS="-DRTB_STREAM_CS=1 -DRTB_TABLE_EL=1" python -m riptable_b200.build --force > /dev/null 2>&1; run cs_el
RTB_NVCC_FLAGS="-DRTB_TABLE_EL=1" python -m riptable_b200.build --force > /dev/null 2>&1; run el
