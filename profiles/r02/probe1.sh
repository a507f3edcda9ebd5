#!/bin/bash
# round 2, GPU call 1: new parity tests, the fused step against the two-call step, table-load and occupancy variants
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_pack_sort.py -x -q > gpurun_out/r2_t1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_t1.log
tail -5 gpurun_out/r2_t1.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
RTB_GB_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_trace.json 2> gpurun_out/r2_bench_trace.err
for load in 0.5 0.35; do
  RTB_GB_LOAD=$load timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_load$load.json 2> gpurun_out/r2_bench_load$load.err
done
RTB_GB_CTAS=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_ctas3.json 2>/dev/null
RTB_NVCC_FLAGS=-DRTB_FUSE_MINB=4 python -m riptable_b200.build --force > gpurun_out/r2_build4.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_minb4.json 2> gpurun_out/r2_bench_minb4.err
for f in gpurun_out/r2_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print("value %.1f G rows/s  ms/step %.3f  two_call %s  e2e %s" % (d["value"]/1e9, d["ms_per_step"], d.get("two_call") and round(d["two_call"]["ms_per_step"],3), d.get("e2e") and d["e2e"].get("value")))
    print({k:(round(v["ms_total"],2), v["launches"], round(v["achieved_gbs"])) for k,v in d["kernels"].items()}, d["clocks"])
except Exception as e:
    print("unreadable", e)
PY
done
