#!/bin/bash
# A/B timing of the kernel variants that sit behind build macros.
#   here (no GPU):    scripts/time_variants.sh build      # side libraries find_tfbs_b200/libtfbs_b200_<name>.so, they travel with gpurun
#   on the B200:      gpurun --timeout 900 -- scripts/time_variants.sh run
# `run` first checks every variant against the oracle on the real hardware (a fast subset of the parity tests), then times the
# default library and every variant with the same short bench command and prints one line each.
set -u
cd "$(dirname "$0")/.."
declare -A VARIANTS=(
  [tile2k]="-DTFBS_TILE_POS=2048 -DTFBS_MAX_PIECES=32"
  [fan512]="-DTFBS_FAN_THREADS=512"
  [fansplit8]="-DTFBS_FAN_SPLIT=8"
  [warps28]="-DTFBS_SCAN_WARPS=28"
  [grab16]="-DTFBS_PER_GRAB=16"
)
case "${1:-}" in
build)
  for n in "${!VARIANTS[@]}"; do
    make -C find_tfbs_b200/csrc variant NAME=$n DEFS="${VARIANTS[$n]}" > /dev/null || { echo "build of $n failed"; exit 1; }
    echo "$n: $(grep -A2 'k_scanILi3' find_tfbs_b200/csrc/build_$n.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'k_scanILi3' find_tfbs_b200/csrc/build_$n.log | grep -o '[0-9]* bytes spill stores' | head -1)"
  done ;;
run)
  mkdir -p gpurun_out
  SUBSET="fixture or gataa or synthetic_small or delta_scoring or pattern_chunks or config5 or threshold or audit"
  for lib in find_tfbs_b200/libtfbs_b200.so find_tfbs_b200/libtfbs_b200_*.so; do
    name=$(basename $lib .so)
    if ! TFBS_B200_LIB=$PWD/$lib timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" > gpurun_out/variant_$name.tests.log 2>&1; then
      echo "$name: PARITY FAILED ($(tail -1 gpurun_out/variant_$name.tests.log))"; continue
    fi
    TFBS_B200_LIB=$PWD/$lib timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-full-scan > gpurun_out/variant_$name.json 2> gpurun_out/variant_$name.err \
      || { echo "$name: bench failed"; continue; }
    python - "$name" gpurun_out/variant_$name.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
s = d["stages_ms"]
print("%-28s step %.3f ms  e2e %.3f ms  k_scan %.3f ms (frac %.3f)  group %.2f build %.2f scan-stage %.2f count %.2f" %
      (sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"],
       s["ms_group"], s["ms_build"], s["ms_scan"], s["ms_count"]))
PY
  done ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
