#!/usr/bin/env bash
# round-2 evidence run: full GPU suite, default bench line, launch list, ncu --set full of one U-ConvBlock's kernels
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rs 2>&1 | tail -25 > gpurun_out/r2z_pytest.txt
tail -2 gpurun_out/r2z_pytest.txt
timeout -k 10 600 python bench.py 2> gpurun_out/r2z_bench.err | tail -1 > gpurun_out/r2z_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2z_bench.json'))
print('bench %.1f mix/s %.3f ms e2e %.1f fwd_hbm %.3f per_block %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac']))
for k in d['roofline']['kernels']: print('  %-70s %.1f us %.3f' % (k['kernel'], k['avg_launch_ms']*1e3, k['frac']))
print(d['roofline']['kernel'], d['roofline']['frac']); print('others', [(o['workload'], round(o['value'],1), round(o['forward_hbm_frac'],3)) for o in d['other_configs']])
print('eager', d['eager_cuda_baseline']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'lat', d['latency_b1']['ms'], 'clocks', d['clocks'])"
timeout -k 10 300 python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2z_launches.csv python tools/profile_forward.py --iters 2 > /dev/null 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"pw_mma_kernel|dw_pyramid|merge_pyramid|pyramid_solve" -s 84 -c 7 \
    -o gpurun_out/prof_r2z_block -f python tools/profile_forward.py --iters 2 > gpurun_out/r2z_ncu.log 2>&1
tail -1 gpurun_out/r2z_ncu.log
