#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "tensor_core or pointwise or tac" 2>&1 | tail -3
  timeout -k 10 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_prepost.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
  for lib in "" variants/mask_tma4.so; do
  echo "=== ${lib:-product}"
  export SDR_B200_LIB=${lib:+$PWD/$lib}
  [ -z "$lib" ] && unset SDR_B200_LIB
  timeout -k 10 200 python tools/bench_stages.py --reps 7 --only mask 2>&1 | grep kernel | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%8.1f us  hbm %.3f  %s' % (d['ms'] * 1000, d['frac_hbm'], d['kernel']))"
  timeout -k 10 200 python tools/bench_stages.py --reps 7 --only mask --workload improved_u36_4096_16k --batch 8 2>&1 | grep kernel | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('cfg5/B8 %8.1f us  hbm %.3f  %s' % (d['ms'] * 1000, d['frac_hbm'], d['kernel']))"
  done
  unset SDR_B200_LIB
  for wl in improved_u16_512 groupcomm_u8_512; do
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-other-configs --workload $wl 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench $wl %.1f mixtures/s  %.3f ms/step  e2e %.1f  fwd_hbm %.3f per_block %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac']))
for k in d['roofline']['kernels']: print('    %-60s %.1f us %.3f' % (k['kernel'], k['avg_launch_ms']*1e3, k['frac']))"
  done
} 2>&1 | tee gpurun_out/r2o.txt
