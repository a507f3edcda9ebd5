#!/bin/bash
# final-form measurements: N=1 full line (e2e + cpu leg), reference arm, extra configurations
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n1","r02_bench_ref"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][0])
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e",{}).get("value"), "cpu", d.get("cpu_baseline"))
    except Exception as e: print(f, "unreadable", e)
PY
timeout 900 python profiles/bench_extra.py > gpurun_out/r02_bench_extra.jsonl 2> gpurun_out/r02_bench_extra.err; echo "extra rc=$?"; tail -30 gpurun_out/r02_bench_extra.jsonl | cut -c1-220
