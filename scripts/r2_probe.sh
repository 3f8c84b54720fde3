#!/bin/bash
# fan-out statistics of configs[2] (variant built with -DTFBS_FAN_STATS) and the per-launch times of a configs[3] step
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2probe}
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0"
TFBS_DEBUG=1 TFBS_B200_LIB=find_tfbs_b200/libtfbs_b200_fanstats.so timeout 300 $B --steps 2 --warmup 3 > $O/fanstats.json 2> $O/fanstats.err; grep "fan-out" $O/fanstats.err | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_c3.csv python bench.py --workload configs3 --c3-regions 8 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c3.log 2>&1
python - "$O" <<'PY'
import sys, os, csv, collections
O = sys.argv[1]
rows = list(csv.DictReader(l for l in open(os.path.join(O, "launches_c3.csv")) if not l.startswith("==")))
names = [r["Kernel Name"].split("(")[0] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
t, c = collections.OrderedDict(), collections.Counter()
for n, v in zip(names, vals):
    t[n] = t.get(n, 0) + v; c[n] += 1
tot = sum(t.values())
print("configs3 all launches: %d, %.3f ms" % (len(names), tot / 1e6))
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:16]:
    print("  %-34s n=%3d %9.3f ms %5.1f%%" % (k[:34], c[k], v / 1e6, 100 * v / tot))
PY
