#!/bin/bash
set -u
mkdir -p gpurun_out
bash profiles/r02/probe14.sh 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_n4.json") if l.startswith("{")][0])
print("N=4 value %.1f G rows/s ms/step %.3f" % (d["value"]/1e9, d["ms_per_step"]), "strong", d["strong"]["ms_per_step"], d["strong"]["value"]/1e9, d["parity_check"]["status"], "e2e", d["e2e"]["value"])
PY
timeout 600 python -m pytest tests/test_gpu_pack_sort.py -x -q -k "lexsort or group_from or glue or cutoffs" > gpurun_out/r02_t_gfl2.log 2>&1; echo "gfl rc=$?"; tail -2 gpurun_out/r02_t_gfl2.log
timeout 900 python profiles/bench_extra.py > gpurun_out/r02_bench_extra3.jsonl 2> gpurun_out/r02_bench_extra3.err; grep "C4" gpurun_out/r02_bench_extra3.jsonl | cut -c1-120
