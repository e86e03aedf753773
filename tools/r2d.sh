#!/usr/bin/env bash
# A/B: the CTA-pair GEMM (product: + lean transform loop) against the round-1 single-CTA kernels
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for lib in "" variants/pair_lean.so; do
  echo "=== ${lib:-product}"
  export SDR_B200_LIB=${lib:+$PWD/$lib}
  [ -z "$lib" ] && unset SDR_B200_LIB
  timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "tensor_core" 2>&1 | tail -1
  timeout -k 10 300 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "golden or cfg2" 2>&1 | tail -1
  timeout -k 10 200 python tools/bench_stages.py --reps 7 --only tcgen05 2>&1 | grep kernel | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%8.1f us  hbm %.3f  %s' % (d['ms'] * 1000, d['frac_hbm'], d['kernel']))"
  for wl in improved_u16_512 improved_u36_2048; do
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-other-configs --workload $wl 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench $wl %.1f mixtures/s  %.3f ms/step  e2e %.1f  fwd_hbm %.3f per_block %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac']))"
  done
done
} 2>&1 | tee gpurun_out/r2d.txt
