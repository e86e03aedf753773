#!/usr/bin/env bash
# original SuDoRM-RF (variant 3): new stage + model tests first, then the rest of the GPU suite, smoke, bench lines
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL="original or pc or residual_norm or softmax_gate"
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "$SEL" -s 2>&1 | tail -60 > gpurun_out/r2s1_new.txt
echo "== new tests"; grep -E "passed|failed|error" gpurun_out/r2s1_new.txt | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2s1_new.txt | head -20
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not ($SEL)" -rs 2>&1 | tail -40 > gpurun_out/r2s1_rest.txt
echo "== rest of the suite"; grep -E "passed|failed|error" gpurun_out/r2s1_rest.txt | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r2s1_rest.txt | head -20
timeout -k 10 400 python bench.py --workload original_u16_512 --no-other-configs 2> gpurun_out/r2s1_bench_orig.err | tail -1 > gpurun_out/r2s1_bench_orig.json
tail -3 gpurun_out/r2s1_bench_orig.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2s1_bench_orig.json'))
    print('original: %.1f mix/s %.3f ms e2e %.1f fwd_hbm %.3f launches %d eager %s' % (
        d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['gpu_launches_per_step'],
        d.get('eager_cuda_baseline', {}).get('value')))
except Exception as ex:
    print('original bench line unreadable:', ex)
PY
timeout -k 10 600 python bench.py 2> gpurun_out/r2s1_bench.err | tail -1 > gpurun_out/r2s1_bench.json
tail -3 gpurun_out/r2s1_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2s1_bench.json'))
    print('bench %.1f mix/s %.3f ms e2e %.1f fwd_hbm %.3f per_block %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['forward_hbm']['frac'], d['roofline']['per_block']['frac']))
    for k in d['roofline']['kernels']: print('  %-70s %.1f us %.3f' % (k['kernel'][:70], k['avg_launch_ms']*1e3, k['frac']))
    print('others', [(o['workload'], round(o['value'], 1), round(o['forward_hbm_frac'], 3)) for o in d['other_configs']])
    print('eager', d['eager_cuda_baseline']['value'], 'cpu', d['cpu_baseline']['value'], 'lat', d['latency_b1']['ms'], 'clocks', d['clocks'])
except Exception as ex:
    print('bench line unreadable:', ex)
PY
