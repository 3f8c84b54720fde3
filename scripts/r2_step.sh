#!/bin/bash
# Round 2 work horse: GPU parity suite, then bench lines of both workloads, then the per-launch device times of one configs[2] step.
#   gpurun --timeout 1500 -- scripts/r2_step.sh <tag> [pytest -k expression]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2x}
SEL=${2:-}
O=gpurun_out/$TAG
mkdir -p $O
if [ -n "$SEL" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/tests.log 2>&1; else timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; fi
tail -5 $O/tests.log
B="python bench.py --no-driver"
timeout 900 $B --cpu-seconds 5 > $O/bench.json 2> $O/bench.err || { echo "bench failed"; tail -5 $O/bench.err; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2.csv $B --no-cpu-baseline --no-secondary --sustain-seconds 0 --steps 2 --warmup 2 > $O/ncu_list_c2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c1.csv $B --workload configs1 --no-cpu-baseline --no-full-scan --sustain-seconds 0 --steps 2 --warmup 2 > $O/ncu_list_c1.log 2>&1
python - "$O" <<'PY'
import json, sys, os, csv, collections
O = sys.argv[1]
try:
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
    print("e2e_dense", d.get("e2e_dense_rows"))
    print("cpu", {k: d.get("cpu_baseline", {}).get(k) for k in ("value", "cores", "regions", "seconds")})
except Exception as e:
    print("bench.json unreadable", e)
for f in ("launches_c2.csv", "launches_c1.csv"):
    try:
        rows = list(csv.DictReader(l for l in open(os.path.join(O, f)) if not l.startswith("==")))
    except Exception as e:
        print(f, "unreadable", e); continue
    names = [r["Kernel Name"].split("(")[0] for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
    idx = [i for i, n in enumerate(names) if "k_status_init" in n]
    if len(idx) < 3: print(f, "no step boundary"); continue
    s0, s1 = idx[-2], idx[-1]
    t, c = collections.OrderedDict(), collections.Counter()
    for n, v in zip(names[s0:s1], vals[s0:s1]):
        t[n] = t.get(n, 0) + v; c[n] += 1
    tot = sum(t.values())
    print(f, "one step: %d launches, %.3f ms" % (s1 - s0, tot / 1e6))
    for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:18]:
        print("  %-34s n=%3d %9.3f ms %5.1f%%" % (k[:34], c[k], v / 1e6, 100 * v / tot))
PY
