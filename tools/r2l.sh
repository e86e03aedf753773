#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "tac or pointwise" 2>&1 | tail -12
  timeout -k 10 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_prepost.py -m gpu -q -p no:cacheprovider -k "groupcomm or cfg4 or golden" 2>&1 | tail -5
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-other-configs --workload groupcomm_u8_512 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench gc %.1f mixtures/s  %.3f ms/step  e2e %.1f  fwd_hbm %.3f per_block %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac']))
for k in d['roofline']['kernels']: print('    %-60s %.1f us %.3f' % (k['kernel'], k['avg_launch_ms']*1e3, k['frac']))"
  timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2l_launches_gc.csv python tools/profile_forward.py --workload groupcomm_u8_512 --iters 2 > /dev/null 2>&1
} 2>&1 | tee gpurun_out/r2l.txt
