#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02_t_all2.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/r02_t_all2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_bench_c.json 2>/dev/null; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_c.json") if l.startswith("{")][0])
print("N=1 value %.1f G rows/s ms/step %.3f two_call %.3f" % (d["value"]/1e9, d["ms_per_step"], d["two_call"]["ms_per_step"]))
PY
python -c "import __graft_entry__ as g; g.smoke()"
