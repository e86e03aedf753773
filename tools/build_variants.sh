#!/usr/bin/env bash
# Builds the A/B variants of libsudormrf_b200.so next to the product build (run HERE, before gpurun: nvcc
# cross-compiles without a GPU, and the .so files travel with the snapshot).  Usage: tools/build_variants.sh [name=flags ...]
# Default set = the round-2 candidates of profiles/r02_plan.md.  Variants land in ./variants/<name>.so (git-ignored).
set -euo pipefail
cd "$(dirname "$0")/.."
ROOT=$PWD
CSRC=$ROOT/sudo_rm_rf_b200/csrc
OUT=$ROOT/variants
mkdir -p "$OUT"
if [ $# -eq 0 ]; then
  set -- "epi8=-DSDR_MMA_EPI_WARPS=8" "lean=-DSDR_MMA_LEAN_PRODUCER=1" \
         "epi8_lean=-DSDR_MMA_EPI_WARPS=8 -DSDR_MMA_LEAN_PRODUCER=1" \
         "dwpipe=-DSDR_DW_PIPELINE=1" "dwchain=-DSDR_DW_CHAIN=1"
fi
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  bdir=$(mktemp -d)
  echo "== $name: $flags"
  for f in api pointwise pointwise_mma levels pyramid causal original frontback tac prepost; do
    nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
         --expt-relaxed-constexpr -Xptxas -v $flags -c "$CSRC/$f.cu" -o "$bdir/$f.o" 2> "$bdir/$f.log" &
  done
  wait
  grep -hE "spill stores" "$bdir"/*.log | sort | uniq -c | sort -rn | head -3
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/$name.so" "$bdir"/*.o -lcuda
  rm -rf "$bdir"
done
ls -la "$OUT"
