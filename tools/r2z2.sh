#!/usr/bin/env bash
# round-2 (second session) evidence run: full GPU suite, smoke, default bench line, reference arm, launch list of the
# default forward, ncu --set full of the kernels added in this session
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rs 2>&1 | tail -25 > gpurun_out/r2z2_pytest.txt
tail -3 gpurun_out/r2z2_pytest.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout -k 10 600 python bench.py 2> gpurun_out/r2z2_bench.err | tail -1 > gpurun_out/r2z2_bench.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2z2_bench.json'))
print('bench %.1f mix/s %.3f ms e2e %.1f fwd_hbm %.3f per_block %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac']))
for k in d['roofline']['kernels']: print('  %-70s %.1f us %.3f' % (k['kernel'][:70], k['avg_launch_ms']*1e3, k['frac']))
print('others', [(o['workload'], round(o['value'], 1), round(o['forward_hbm_frac'], 3)) for o in d['other_configs']])
print('eager', d['eager_cuda_baseline']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'lat', d['latency_b1']['ms'], 'clocks', d['clocks'])
PY
timeout -k 10 300 python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2z2_launches.csv python tools/profile_forward.py --iters 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2z2_launches.csv 2>/dev/null | head -12
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"pw_tile_kernel|tac_mma16" -s 3 -c 3 \
    -o gpurun_out/prof_r2z2_gc -f python tools/profile_forward.py --workload groupcomm_u8_512 --iters 1 > gpurun_out/r2z2_gc.log 2>&1
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"causal_pyramid" -s 2 -c 1 \
    -o gpurun_out/prof_r2z2_causal -f python tools/profile_forward.py --workload causal_u16_512 --iters 1 > gpurun_out/r2z2_causal.log 2>&1
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"softmax_gate|residual_norm|pw_mma_kernel|dw_pyramid" -s 8 -c 8 \
    -o gpurun_out/prof_r2z2_orig -f python tools/profile_forward.py --workload original_u16_512 --iters 1 > gpurun_out/r2z2_orig.log 2>&1
ls -la gpurun_out/prof_r2z2_*.ncu-rep
