#!/bin/bash
# the default library and every find_tfbs_b200/libtfbs_b200_<name>.so on the same short configs[2] bench, at full size and at 1/8
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2v}
mkdir -p $O
B="python bench.py --no-driver --no-cpu-baseline --no-secondary --sustain-seconds 0 --steps 10 --warmup 2"
for lib in find_tfbs_b200/libtfbs_b200.so find_tfbs_b200/libtfbs_b200_*.so; do
  name=$(basename $lib .so)
  for sc in 1.0 0.125; do
  TFBS_B200_LIB=$PWD/$lib timeout 300 $B --scale $sc > $O/$name.$sc.json 2> $O/$name.err || { echo "$name failed"; tail -2 $O/$name.err; continue; }
  python - $name $sc $O/$name.$sc.json <<'PY'
import json, sys
d = json.load(open(sys.argv[3])); s = d["rank0"]["stages_ms"]
print("%-24s scale %-6s step %7.3f e2e %7.3f k_scan %6.3f | group %.2f build %.2f scan-stage %.2f count %.2f" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_step"], s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"]))
PY
  done
done
