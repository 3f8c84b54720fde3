#!/bin/bash
# ncu --set full (with source counters) of the bookkeeping kernels of one configs[2] step (scale 0.25): scripts/r2_ncu_many.sh <tag>
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2many}
K='k_walk|k_signatures|k_cfg_runs|k_fanout|k_members|k_group_insert|k_redirect|k_cfg_resolve'
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0 --scale 0.25 --steps 2 --warmup 2"
# 9 matching launches per step (k_members twice); skip the first two steps
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip 18 -c 9 -o $O/many -f $B > $O/ncu.log 2>&1
ncu -i $O/many.ncu-rep --page raw --csv > $O/many_raw.csv 2> /dev/null
ls -la $O/many.ncu-rep
tail -2 $O/ncu.log
