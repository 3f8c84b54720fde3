#!/bin/bash
# ncu --set full of one kernel of a configs[2] step (scale 0.25), with the source page: scripts/r2_ncu_one.sh <tag> <kernel regex>
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2k}
K=${2:-k_fanout}
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0 --scale 0.25 --steps 2 --warmup 2"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip 2 -c 1 -o $O/one -f $B > $O/ncu.log 2>&1
ncu -i $O/one.ncu-rep --page raw --csv > $O/one_raw.csv 2> /dev/null
ncu -i $O/one.ncu-rep --page source --csv > $O/one_source.csv 2> /dev/null
ncu -i $O/one.ncu-rep --page details > $O/one_details.txt 2> /dev/null
ls -la $O/one.ncu-rep
tail -2 $O/ncu.log
