#!/bin/bash
# `ncu --set full` of the kernels that make up a configs[2] step (scale 0.25: same per-region shape, a quarter of the regions), plus
# shared-memory table budgets of the scan and the driver's wall time on the file sets.
#   gpurun --timeout 1800 -- scripts/r2_ncu.sh <tag>
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2n}
O=gpurun_out/$TAG
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0"
timeout 300 $B --steps 5 --warmup 2 > $O/bench_default.json 2> $O/bench_default.err || tail -3 $O/bench_default.err
for kb in 128 160 180; do
  timeout 300 $B --steps 5 --warmup 2 --option table_budget_kb=$kb > $O/bench_tb$kb.json 2> $O/bench_tb$kb.err || tail -3 $O/bench_tb$kb.err
done
python - "$O" <<'PY'
import json, sys, os, glob
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench_*.json"))):
    try:
        d = json.load(open(f)); s = d["rank0"]["stages_ms"]; rf = d["roofline"]
        print("%-22s step %7.3f e2e %7.3f k_scan %6.3f (%d launches) frac %.3f | group %.2f build %.2f scan-stage %.2f count %.2f" %
              (os.path.basename(f), d["ms_per_step"], d["e2e"]["ms_per_step"], rf["ms_per_step"], rf["launches_per_step"], rf["frac"], s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
K='regex:k_fanout|k_scan|k_walk|k_group_lookup|k_signatures|k_group_insert|k_cfg_runs|k_members|k_seq_insert|k_nominal|k_cfg_resolve|k_row_headers'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 34 -c 17 -o $O/full_c2 -f $B --scale 0.25 --steps 2 --warmup 2 > $O/ncu_full.log 2>&1
ncu -i $O/full_c2.ncu-rep --page raw --csv > $O/full_c2_raw.csv 2> /dev/null
ncu -i $O/full_c2.ncu-rep --page source --csv -k regex:k_fanout > $O/fanout_source.csv 2> /dev/null
rm -f $O/full_c2.ncu-rep
tail -3 $O/ncu_full.log
for name in cfg1 cfg2; do
  [ -f scratch_data/$name/args.txt ] || continue
  for t in 16 64; do
    T0=$(date +%s.%N)
    find_tfbs_b200/find-tfbs-b200 $(cat scratch_data/$name/args.txt) --output /tmp/out_$name.vcf.gz --threads $t > $O/driver_${name}_t$t.log 2>&1
    T1=$(date +%s.%N)
    echo "$name threads=$t: $(echo "$T1 - $T0" | bc) s wall | $(tail -3 $O/driver_${name}_t$t.log | head -2 | tr '\n' ' ')"
  done
  ls -la /tmp/out_$name.vcf.gz | awk '{print $5 " bytes"}'
done
nproc
