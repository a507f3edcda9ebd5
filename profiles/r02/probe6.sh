#!/bin/bash
# 2 GPUs: NCCL parity tests of the sharded paths, then the bench at N=2 (weak + strong + parity legs)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist2.py tests/test_gpu_dist.py -x -q > gpurun_out/r2_t4_dist.log 2>&1; echo "dist rc=$?"; tail -15 gpurun_out/r2_t4_dist.log
RTB_DIST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench2 rc=$?"
grep -h "rtb dist r1" gpurun_out/r2_bench_n2.json gpurun_out/r2_bench_n2.err | tail -4
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n2.json") if l.startswith("{")][0])
    print("N=2 value %.1f G rows/s ms/step %.3f" % (d["value"]/1e9, d["ms_per_step"]))
    print("strong", d.get("strong")); print("parity", d.get("parity_check"))
except Exception as e:
    print("unreadable", e)
PY
tail -5 gpurun_out/r2_bench_n2.err
