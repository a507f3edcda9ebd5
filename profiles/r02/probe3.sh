#!/bin/bash
set -u
mkdir -p gpurun_out
RTB_GB_TRACE=1 timeout 300 python profiles/r02/dbg_fused.py > gpurun_out/r2_dbg_fused.log 2>&1; tail -30 gpurun_out/r2_dbg_fused.log
for pf in 0 1; do
  RTB_GB_FUSEPF=$pf timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_pf$pf.json 2> gpurun_out/r2_bench_pf$pf.err
done
RTB_GB_FUSEPF=0 RTB_GB_CTAS=8 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_pf0c8.json 2>/dev/null
for f in gpurun_out/r2_bench_pf*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print("value %.1f G rows/s  ms/step %.3f  two_call %s" % (d["value"]/1e9, d["ms_per_step"], d.get("two_call") and (round(d["two_call"]["ms_per_step"],3), round(d["two_call"]["hash_ms_per_step"],3))))
    print({k:(round(v["ms_total"],2), v["launches"], round(v["achieved_gbs"])) for k,v in d["kernels"].items()})
except Exception as e:
    print("unreadable", e)
PY
done
