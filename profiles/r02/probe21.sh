#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_groupby.py -x -q -k "pipeline or fused or stream" > gpurun_out/r02_t_pipe.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_t_pipe.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_d.json") if l.startswith("{")][0])
print("N=1 value %.1f G rows/s ms/step %.3f" % (d["value"]/1e9, d["ms_per_step"]), "e2e", d["e2e"], "cpu", d["cpu_baseline"].get("parity"), d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/r02_bench_d.err
