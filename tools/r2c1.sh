#!/usr/bin/env bash
# causal variant: stage + model parity, smoke, and a bench line of the causal workload
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "causal" -s 2>&1 | tail -40 > gpurun_out/r2c1_pytest.txt
tail -30 gpurun_out/r2c1_pytest.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout -k 10 600 python bench.py --workload causal_u16_512 --no-other-configs 2> gpurun_out/r2c1_bench.err | tail -1 > gpurun_out/r2c1_bench.json
tail -5 gpurun_out/r2c1_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c1_bench.json'))
print('bench %.1f mix/s %.3f ms e2e %.1f fwd_hbm %.3f per_block %.3f (%.1f us)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac'], d['roofline']['per_block']['ms']*1e3))
for k in d['roofline']['kernels']: print('  %-90s %.1f us %.3f' % (k['kernel'][:90], k['avg_launch_ms']*1e3, k['frac']))
print('eager', d['eager_cuda_baseline']['value'], 'cpu', d['cpu_baseline']['value'], 'lat', d['latency_b1']['ms'], 'launches', d['gpu_launches'])"
