#!/bin/bash
# N GPUs of one box: the 2-rank tests, then bench.py at N = 1 and N = given (torchrun, one rank per GPU), reference arm skipped.
#   gpurun --gpus N --timeout 1500 -- scripts/r2_multi.sh <tag> N
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-r2multi}
N=${2:-2}
mkdir -p $O
nvidia-smi -L | head -8
[ "${3:-}" = "notests" ] || timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err || tail -3 $O/bench_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err || tail -5 $O/bench_n$n.err
done
python - "$O" <<'PY'
import json, sys, os, glob
base = None
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench_n*.json")), key=lambda p: int(p.split("_n")[1].split(".")[0])):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    n = d["n_gpus"]
    if n == 1: base = d
    eff = (d["value"] / base["value"] / n, d["e2e"]["value"] / base["e2e"]["value"] / n) if base else (None, None)
    w = d.get("wall_s", {})
    print("N=%d step %7.3f ms value %.3e (eff %s) | e2e %7.3f ms %.3e (eff %s) h2d %.1f MB d2h %.1f MB | gathered %s | driver wall %s s" %
          (n, d["ms_per_step"], d["value"], "%.3f" % eff[0] if eff[0] else "-", d["e2e"]["ms_per_step"], d["e2e"]["value"], "%.3f" % eff[1] if eff[1] else "-",
           d["e2e"]["h2d_bytes_per_step"] / 1e6, d["e2e"]["d2h_bytes_per_step"] / 1e6, d["e2e"]["gathered_on_rank0"], w.get("value", w)))
    if "phases" in w: print("   ", w["phases"][-2][:400])
PY
