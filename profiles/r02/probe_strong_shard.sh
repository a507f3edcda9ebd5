#!/bin/bash
# What an N=8 strong-scaling shard (125 M rows per rank) costs per phase, measured on 2 GPUs; seeded vs generic path.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
ARGS="--rows 125000000 --steps 6 --warmup 3 --no-e2e --no-cpu --no-strong --no-parity"
echo "== N=1, 125 M rows"; python bench.py --gpus 1 $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
echo "== N=2 seeded"; RTB_DIST_TRACE=1 $TR bench.py --gpus 2 $ARGS 2>&1 | grep -v "^\[W\|warn" | tail -8 | cut -c1-400
echo "== N=2 generic"; RTB_DIST_SEEDED=0 RTB_DIST_TRACE=1 $TR bench.py --gpus 2 $ARGS 2>&1 | grep -v "^\[W\|warn" | tail -8 | cut -c1-400
