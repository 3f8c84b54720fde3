#!/bin/bash
# what one rank of an N = 8 run sees: the configs[2] block at 1/8 of the regions on one GPU, with the per-launch device times
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2sp}
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0"
for sc in 1.0 0.5 0.25 0.125; do
  timeout 300 $B --steps 20 --warmup 3 --scale $sc > $O/bench_$sc.json 2> $O/bench_$sc.err || tail -3 $O/bench_$sc.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv $B --scale 0.125 --steps 2 --warmup 2 > $O/ncu_list.log 2>&1
python - "$O" <<'PY'
import json, sys, os, csv, collections
O = sys.argv[1]
for sc in ("1.0", "0.5", "0.25", "0.125"):
    try:
        d = json.load(open(os.path.join(O, "bench_%s.json" % sc))); s = d["rank0"]["stages_ms"]
        print("scale %-6s regions %5d step %7.3f e2e %7.3f (per region %.2f us) k_scan %6.3f | group %.2f build %.2f scan-stage %.2f count %.2f | launches %d" %
              (sc, d["config"]["regions"], d["ms_per_step"], d["e2e"]["ms_per_step"], 1e3 * d["ms_per_step"] / d["config"]["regions"], d["roofline"]["ms_per_step"], s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"], d["rank0"]["launches_per_step"]))
    except Exception as e:
        print(sc, "unreadable", e)
rows = list(csv.DictReader(l for l in open(os.path.join(O, "launches.csv")) if not l.startswith("==")))
names = [r["Kernel Name"].split("(")[0] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
idx = [i for i, n in enumerate(names) if "k_status_init" in n]
s0, s1 = idx[-2], idx[-1]
t, c = collections.OrderedDict(), collections.Counter()
for n, v in zip(names[s0:s1], vals[s0:s1]):
    t[n] = t.get(n, 0) + v; c[n] += 1
tot = sum(t.values())
print("scale 0.125 one step: %d launches, %.3f ms of kernel time" % (s1 - s0, tot / 1e6))
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:40]:
    print("  %-34s n=%3d %9.3f ms %5.1f%%" % (k[:34], c[k], v / 1e6, 100 * v / tot))
PY
