#!/bin/bash
# Final single-GPU evidence of round 2: the whole GPU suite, the bench line, the reference arm, the ncu launch list, the
# other BASELINE configurations.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02f_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r02f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu_l.log 2>&1
echo "launch list rc=$?"
timeout 900 python profiles/bench_extra.py > gpurun_out/r02f_bench_extra.jsonl 2> gpurun_out/r02f_bench_extra.err; echo "extra rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02f_bench_n1.json") if l.startswith("{")][0])
print("N=1 value %.2f G rows/s, %.3f ms/step, e2e %.3f G, frac %.3f, cpu %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["roofline"]["frac"], d["cpu_baseline"]["value"]))
r=json.loads([l for l in open("gpurun_out/r02f_bench_ref.json") if l.startswith("{")][0])
print("reference arm", r["value"], r.get("cpu_baseline"))
for l in open("gpurun_out/r02f_bench_extra.jsonl"):
    if l.startswith("{"):
        x=json.loads(l); print("%-55s %9.2f ms" % (x.get("op"), x.get("ms", 0)))
PY
