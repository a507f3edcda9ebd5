#!/bin/bash
set -u
mkdir -p gpurun_out
G=${1:-2}
timeout 600 python -m pytest tests/test_gpu_dist2.py -x -q > gpurun_out/r02_t_dist_n$G.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/r02_t_dist_n$G.log
RTB_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 3 --warmup 3 --no-e2e --no-strong --no-parity > gpurun_out/r02_trace_n$G.log 2> gpurun_out/r02_trace_n$G.err
grep -h "rtb dist r0\]\|rtb dist r1\]" gpurun_out/r02_trace_n$G.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/r02_bench_n$G.json 2> gpurun_out/r02_bench_n$G.err; echo "bench rc=$?"
python - $G <<'PY'
import json,sys
G=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r02_bench_n{G}.json") if l.startswith("{")][0])
    print(f"N={G} value %.1f G rows/s ms/step %.3f e2e %s" % (d["value"]/1e9, d["ms_per_step"], d.get("e2e") and d["e2e"].get("value")))
    print("strong", d.get("strong")); print("parity", d.get("parity_check"))
except Exception as e:
    print("unreadable", e)
PY
