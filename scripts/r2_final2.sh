#!/bin/bash
# Evidence of the final state of round 2 (second session): the whole GPU test suite, the bench line exactly as the driver runs it
# (plus the reference arm), the per-launch device times of the same command, `ncu --set full` of the scan kernel on both workloads
# and of the fan-out (with per-source-line instruction tables), and the configs[3] line.
#   gpurun --timeout 2700 -- scripts/r2_final2.sh <tag>
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2u}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; tail -4 $O/tests.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err || tail -5 $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err || tail -5 $O/bench_reference.err
timeout 900 python bench.py --workload configs3 --steps 3 --warmup 3 --cpu-seconds 10 > $O/bench_c3.json 2> $O/bench_c3.err || tail -5 $O/bench_c3.err
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0 --steps 2 --warmup 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_c2.csv $B > $O/ncu_list_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_c1.csv $B --workload configs1 --no-full-scan > $O/ncu_list_c1.log 2>&1
# k_scan: the three launches of a steady-state resident step of configs[2] (4 warm-up blocks x 3 launches skipped), one launch of configs[1]
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan --launch-skip 12 -c 3 -o $O/scan_c2 -f $B > $O/ncu_scan_c2.log 2>&1
ncu -i $O/scan_c2.ncu-rep --page raw --csv > $O/scan_c2_raw.csv 2> /dev/null
ncu -i $O/scan_c2.ncu-rep --page source --csv > $O/scan_c2_source.csv 2> /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scan --launch-skip 4 -c 1 -o $O/scan_c1 -f $B --workload configs1 --no-full-scan > $O/ncu_scan_c1.log 2>&1
ncu -i $O/scan_c1.ncu-rep --page raw --csv > $O/scan_c1_raw.csv 2> /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fanout|k_walk" --launch-skip 8 -c 2 -o $O/book_c2 -f $B > $O/ncu_book_c2.log 2>&1
ncu -i $O/book_c2.ncu-rep --page raw --csv > $O/book_c2_raw.csv 2> /dev/null
ncu -i $O/book_c2.ncu-rep --page source --csv > $O/book_c2_source.csv 2> /dev/null
# per-source-line instruction tables (SASS table of ncu x line info of the shipped cubin)
mkdir -p /tmp/sass && (cd /tmp/sass && rm -f *.cubin && cuobjdump -xelf all $OLDPWD/find_tfbs_b200/libtfbs_b200.so > /dev/null 2>&1 && nvdisasm -g -c tfbs.sm_100a.cubin > tfbs.txt)
python scripts/sass_by_line.py $O/book_c2_source.csv /tmp/sass/tfbs.txt 8k_fanoutE 30 "tfbs::k_fanout" > $O/k_fanout_by_line.txt 2>&1
python scripts/sass_by_line.py $O/book_c2_source.csv /tmp/sass/tfbs.txt 6k_walkE 30 "tfbs::k_walk" > $O/k_walk_by_line.txt 2>&1
python scripts/sass_by_line.py $O/scan_c2_source.csv /tmp/sass/tfbs.txt 6k_scanILi3E 30 "k_scan" > $O/k_scan_by_line.txt 2>&1
cuobjdump -sass -fun '_ZN4tfbs6k_scanILi3EEEvNS_8DevBlockENS_7DevSeqsENS_11DevPatternsENS_9DevCountsENS_10DevMatchesENS_10DevRefHitsENS_10DevConfigsEiPKjPKyjPNS_9DevStatusEj' find_tfbs_b200/libtfbs_b200.so 2> /dev/null | grep -E '^\s+/\*[0-9a-f]{4}\*/' | awk '{print $2}' | sed 's/;//' | sort | uniq -c | sort -rn > $O/k_scan_sm100a_opcodes.txt
rm -f $O/*.ncu-rep $O/*_source.csv
python - "$O" <<'PY'
import json, sys, os
O = sys.argv[1]
d = json.load(open(os.path.join(O, "bench.json")))
for name, x in (("configs2", d), ("configs1", d.get("secondary", {}))):
    if "ms_per_step" not in x:
        print(name, x); continue
    s = x["rank0"]["stages_ms"] if "rank0" in x else x["stages_ms"]
    rf = x["roofline"]
    print("%-9s step %8.3f e2e %8.3f (d2h %.1f MB) k_scan %7.3f frac %.3f full %s | group %.2f build %.2f scan-stage %.2f count %.2f | rows %d" %
          (name, x["ms_per_step"], x["e2e"]["ms_per_step"], x["e2e"]["d2h_bytes_per_step"] / 1e6, rf["ms_per_step"], rf["frac"],
           ("%.3f" % rf["full_scan"]["frac"]) if rf.get("full_scan") else "-", s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"], x["rows_per_step"]))
print("sustained", d.get("sustained"))
print("wall_s", d.get("wall_s"))
print("cpu", {k: d.get("cpu_baseline", {}).get(k) for k in ("value", "cores", "regions", "seconds", "chunk")})
r = json.load(open(os.path.join(O, "bench_reference.json")))
print("reference arm value %.3e ms/step %.1f -> e2e ratio %.0f" % (r["value"], r["ms_per_step"], d["e2e"]["value"] / r["value"]))
c = json.load(open(os.path.join(O, "bench_c3.json")))
print("configs3 step %.1f e2e %.1f value %.3e" % (c["ms_per_step"], c["e2e"]["ms_per_step"], c["value"]), c["stages_ms_summed_over_blocks"], "k_scan frac %.3f k1 frac %.4f" % (c["roofline"]["frac"], c["roofline_k1"]["frac"]))
PY
