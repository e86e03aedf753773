#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "pyramid" 2>&1 | tail -5
  timeout -k 10 300 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "golden or cfg2" 2>&1 | tail -3
  for wl in improved_u16_512 improved_u36_2048 groupcomm_u8_512; do
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-other-configs --workload $wl 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench $wl %.1f mixtures/s  %.3f ms/step  e2e %.1f  fwd_hbm %.3f per_block %.3f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac'], d['gpu_launches_per_step']))
for k in d['roofline']['kernels']: print('    %-70s %.1f us %.3f' % (k['kernel'], k['avg_launch_ms']*1e3, k['frac']))"
  done
} 2>&1 | tee gpurun_out/r2r.txt
