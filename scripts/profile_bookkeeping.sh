#!/bin/bash
# First GPU call of a tuning session: where the non-scan 5 ms of a configs[1] step go.
# Runs the bench once without a profiler (the step must exit 0 first), then one launch list and one `ncu --set full` capture per
# bookkeeping kernel plus the scan in both modes.  Everything lands in gpurun_out/; copy what is worth citing into profiles/.
#   gpurun --timeout 900 -- scripts/profile_bookkeeping.sh
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-full-scan"
timeout 120 $CMD > gpurun_out/pb_plain.json 2> gpurun_out/pb_plain.err || { echo "bench failed"; tail -5 gpurun_out/pb_plain.err; exit 1; }
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pb_launches.csv $CMD > gpurun_out/pb_ncu_list.log 2>&1
for k in k_group_finish k_walk k_items k_item_resolve k_emit_list k_scan; do
  # -s 1: skip the warm-up step's launch of the kernel (k_items has two instantiations per step: skip 2, take the <true> one)
  skip=1; [ "$k" = k_items ] && skip=3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s $skip -c 1 -o gpurun_out/pb_$k -f $CMD > gpurun_out/pb_ncu_$k.log 2>&1
  ncu -i gpurun_out/pb_$k.ncu-rep --page raw --csv > gpurun_out/pb_${k}_raw.csv 2>/dev/null
done
# the scan doing the reference's full work (delta scoring off), for the roofline of the kernel itself
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^k_scan" -s 1 -c 1 -o gpurun_out/pb_k_scan_full -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-full-scan --option delta=0 > gpurun_out/pb_ncu_k_scan_full.log 2>&1
ncu -i gpurun_out/pb_k_scan_full.ncu-rep --page raw --csv > gpurun_out/pb_k_scan_full_raw.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(l for l in open("gpurun_out/pb_launches.csv") if not l.startswith("==")))
t = collections.OrderedDict()
for r in rows:
    name = r.get("Kernel Name", "").split("(")[0]
    try:
        v = float(r.get("Metric Value", "0").replace(",", ""))
    except ValueError:
        continue
    t[name] = t.get(name, 0.0) + v
tot = sum(t.values()) or 1.0
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:25]:
    print("%-40s %10.3f ms %5.1f%%" % (k[:40], v / 1e6, 100 * v / tot))
PY
