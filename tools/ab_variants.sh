#!/usr/bin/env bash
# On the GPU box (under gpurun): correctness subset + per-kernel timings for the product and every ./variants/*.so.
# Usage: tools/ab_variants.sh [pytest -k expression] ; results in gpurun_out/ab_variants.txt
set -uo pipefail
cd "$(dirname "$0")/.."
KEXPR=${1:-"tensor_core or depthwise or merge"}
mkdir -p gpurun_out
{
for lib in "" variants/*.so; do
  [ -n "$lib" ] && [ ! -f "$lib" ] && continue
  echo "=== ${lib:-product}"
  export SDR_B200_LIB=${lib:+$PWD/$lib}
  [ -z "$lib" ] && unset SDR_B200_LIB
  timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "$KEXPR" 2>&1 | tail -1
  timeout -k 10 200 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "golden or cfg2" 2>&1 | tail -1
  timeout -k 10 200 python tools/bench_stages.py --reps 5 2>&1 | grep kernel | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'ffma' in d['kernel']: continue
    print('%8.1f us  hbm %.3f  %s' % (d['ms'] * 1000, d['frac_hbm'], d['kernel']))"
  timeout -k 10 200 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench %.1f mixtures/s  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
} 2>&1 | tee gpurun_out/ab_variants.txt
