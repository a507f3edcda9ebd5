#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_groupby.py -x -q -k "small_inputs or 16k_bins or single_key or special or docstring or fused or ismember" > gpurun_out/r2_t5.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2_t5.log
RTB_GB_SMALL=0 timeout 300 python profiles/r02/small_latency.py > gpurun_out/r2_small0.json 2>&1; cat gpurun_out/r2_small0.json
RTB_GB_SMALL=1 timeout 300 python profiles/r02/small_latency.py > gpurun_out/r2_small1.json 2>&1; cat gpurun_out/r2_small1.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_b.json") if l.startswith("{")][0])
print("N=1 value %.1f G rows/s ms/step %.3f two_call %.3f" % (d["value"]/1e9, d["ms_per_step"], d["two_call"]["ms_per_step"]))
PY
