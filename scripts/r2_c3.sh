#!/bin/bash
# configs[3] (sample blocks): parity case first, then the bench object at two sizes
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2c3}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q -x -k "config4 or sample_block" > $O/tests.log 2>&1; tail -3 $O/tests.log
timeout 400 python bench.py --workload configs3 --c3-regions 8 --steps 2 --warmup 3 --cpu-seconds 5 > $O/bench_c3_r8.json 2> $O/bench_c3_r8.err || tail -5 $O/bench_c3_r8.err
timeout 900 python bench.py --workload configs3 --c3-regions ${2:-32} --steps 3 --warmup 3 --cpu-seconds 15 > $O/bench_c3.json 2> $O/bench_c3.err || tail -5 $O/bench_c3.err
python - "$O" <<'PY'
import json, sys, os
for f in ("bench_c3_r8.json", "bench_c3.json"):
    try:
        d = json.load(open(os.path.join(sys.argv[1], f)))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "step %.1f ms e2e %.1f ms value %.3e" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]), d["stages_ms_summed_over_blocks"],
          "k_scan frac %.3f" % d["roofline"]["frac"], "k1 frac %.4f" % d["roofline_k1"]["frac"], "rows", d["e2e"]["rows_all_keys"], d["e2e"]["rows_kept"],
          "groups", d["groups"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
