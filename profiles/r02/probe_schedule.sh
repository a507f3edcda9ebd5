#!/bin/bash
# Round-schedule knobs on the fused C2 call (total ms of the third call), 125 M and 1 B rows.
for rows in 125000000 1000000000; do
  for env in "" "RTB_GB_RATIO=8" "RTB_GB_ROUND0=2097152" "RTB_GB_ROUND0=2097152 RTB_GB_RATIO=8" "RTB_GB_ROUND0=4194304 RTB_GB_RATIO=8" "RTB_GB_FINISH=1073741824"; do
    echo "== rows $rows $env"
    env $env python profiles/r02/trace_c2.py $rows 2>&1 | awk '/---- call 2/{p=1} p' | cut -c1-150
  done
done
