#!/bin/bash
# quick look: bench line of configs[2] (no secondary), per-launch times, heavy-key counter
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2q}
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0 ${EXTRA:-}"
timeout 300 $B --steps 10 --warmup 2 > $O/bench.json 2> $O/bench.err || tail -3 $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_c2.csv $B --steps 2 --warmup 2 > $O/ncu_list_c2.log 2>&1
python - "$O" <<'PY'
import json, sys, os, csv, collections
O = sys.argv[1]
d = json.load(open(os.path.join(O, "bench.json")))
s = d["rank0"]["stages_ms"]; rf = d["roofline"]
print("configs2 step %.3f e2e %.3f k_scan %.3f frac %.3f | group %.2f build %.2f scan-stage %.2f count %.2f | rows %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], rf["ms_per_step"], rf["frac"], s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"], d["rows_per_step"]))
print("rank0", d["rank0"])
rows = list(csv.DictReader(l for l in open(os.path.join(O, "launches_c2.csv")) if not l.startswith("==")))
names = [r["Kernel Name"].split("(")[0] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
idx = [i for i, n in enumerate(names) if "k_status_init" in n]
s0, s1 = idx[-2], idx[-1]
t, c = collections.OrderedDict(), collections.Counter()
for n, v in zip(names[s0:s1], vals[s0:s1]):
    t[n] = t.get(n, 0) + v; c[n] += 1
tot = sum(t.values())
print("one step: %d launches, %.3f ms" % (s1 - s0, tot / 1e6))
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:14]:
    print("  %-34s n=%3d %9.3f ms %5.1f%%" % (k[:34], c[k], v / 1e6, 100 * v / tot))
PY
