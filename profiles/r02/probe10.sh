#!/bin/bash
# ncu evidence for the round-2 bench step: launch list (gpu__time_duration) and a full capture of the fused steady-state kernel
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_l.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gb_round_direct -s 8 -c 3 -o gpurun_out/prof_r02 $CMD > gpurun_out/r02_ncu_f.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/prof_r02.ncu-rep gpurun_out/r02_launches.csv
tail -3 gpurun_out/r02_ncu_f.log
