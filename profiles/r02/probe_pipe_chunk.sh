#!/bin/bash
# Host-mode (e2e) chunk size of the pipelined fused call, 1 B rows on one GPU.
for c in 33554432 16777216 8388608 4194304; do
  echo "== RTB_PIPE_CHUNK_ROWS=$c"
  RTB_PIPE_CHUNK_ROWS=$c python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.2f G  e2e %.3f G rows/s'%(d['value']/1e9, d['e2e']['value']/1e9))"
done
