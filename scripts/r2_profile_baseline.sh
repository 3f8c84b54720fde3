#!/bin/bash
# Round 2, first GPU call: where a step goes on BOTH benchmarked workloads before anything is changed.
#   gpurun --timeout 1500 -- scripts/r2_profile_baseline.sh
# 1. bench lines without a profiler (configs[1], configs[2]); 2. per-launch device times of one configs[2] step;
# 3. `ncu --set full` of every kernel of a configs[2] step at scale 0.25 (same per-region shape, quarter of the regions) and of the
# bookkeeping kernels of a configs[1] step; 4. the parked k_scan / k_group_finish variants, timed on both workloads.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r2a
mkdir -p $O
B1="python bench.py --workload configs1 --no-cpu-baseline --no-full-scan"
B2="python bench.py --workload configs2 --no-cpu-baseline --no-full-scan"
timeout 300 $B1 --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err || { echo "configs1 bench failed"; tail -5 $O/bench_c1.err; }
timeout 600 $B2 --steps 3 --warmup 1 > $O/bench_c2.json 2> $O/bench_c2.err || { echo "configs2 bench failed"; tail -5 $O/bench_c2.err; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c2.csv $B2 --steps 1 --warmup 1 > $O/ncu_list_c2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c1.csv $B1 --steps 1 --warmup 1 > $O/ncu_list_c1.log 2>&1
K='regex:k_signatures|k_group_|k_walk|k_seq_|k_items|k_item_|k_emit_list|k_scan|k_rows_|k_redirect|k_variant_prep|k_ref_prefix'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -c 40 -o $O/full_c2 -f $B2 --scale 0.25 --steps 1 --warmup 0 > $O/ncu_full_c2.log 2>&1
ncu -i $O/full_c2.ncu-rep --page raw --csv > $O/full_c2_raw.csv 2> /dev/null
K1='regex:k_group_finish|k_walk|k_items|k_item_resolve|k_emit_list|k_scan'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K1" -c 8 -o $O/full_c1 -f $B1 --steps 1 --warmup 0 > $O/ncu_full_c1.log 2>&1
ncu -i $O/full_c1.ncu-rep --page raw --csv > $O/full_c1_raw.csv 2> /dev/null
rm -f $O/full_c2.ncu-rep $O/full_c1.ncu-rep  # keep the CSV extracts (the reports exceed what travels back)
for lib in find_tfbs_b200/libtfbs_b200.so find_tfbs_b200/libtfbs_b200_*.so; do
  name=$(basename $lib .so)
  TFBS_B200_LIB=$PWD/$lib timeout 200 $B1 --steps 30 --warmup 3 > $O/var_c1_$name.json 2> $O/var_c1_$name.err || echo "$name configs1 failed"
  TFBS_B200_LIB=$PWD/$lib timeout 300 $B2 --steps 3 --warmup 1 > $O/var_c2_$name.json 2> $O/var_c2_$name.err || echo "$name configs2 failed"
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r2a/var_c*_*.json")) + ["gpurun_out/r2a/bench_c1.json", "gpurun_out/r2a/bench_c2.json"]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(os.path.basename(f), "unreadable", e); continue
    s = d["stages_ms"]
    print("%-40s step %8.3f e2e %8.3f k_scan %7.3f frac %.3f | group %.2f build %.2f scan-stage %.2f count %.2f" %
          (os.path.basename(f), d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"],
           s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"]))
PY
