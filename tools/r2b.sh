#!/usr/bin/env bash
# round-2 GPU call B: first run of the CTA-pair (cta_group::2) GEMM
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "=== stage tests (tensor core)"
timeout -k 10 240 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -x -k "tensor_core" 2>&1 | tail -15
echo "=== model tests"
timeout -k 10 400 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -x -k "golden or cfg2 or cfg3 or cfg5 or cfg1" 2>&1 | tail -15
echo "=== stages"
timeout -k 10 200 python tools/bench_stages.py --reps 5 2>&1 | tee gpurun_out/r2b_stages_raw.txt | grep kernel | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'ffma' in d['kernel']: continue
    print('%8.1f us  hbm %.3f  %s' % (d['ms'] * 1000, d['frac_hbm'], d['kernel']))"
echo "=== bench"
timeout -k 10 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/r2b_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json'))
print('bench %.1f mix/s %.3f ms e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('per_block', d['roofline']['per_block'])
for k in d['roofline']['kernels']: print('  %-60s %.1f us %.3f' % (k['kernel'], k['avg_launch_ms']*1e3, k['frac']))
print('eager', d['eager_cuda_baseline']); print('others', d['other_configs']); print('lat', d['latency_b1']); print('cpu', d['cpu_baseline'])
"
} 2>&1 | tee gpurun_out/r2b.txt
