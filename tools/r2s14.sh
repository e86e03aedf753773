#!/usr/bin/env bash
# TAC: hidden layer evaluated once per group (product) vs twice (variants/tac2pass.so)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "tac or groupcomm or golden or cfg4 or corpus" 2>&1 | tail -2
for v in product tac2pass product tac2pass; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv -k regex:"tac_mma16" \
        --log-file gpurun_out/r2s14_$v.csv python tools/profile_forward.py --workload groupcomm_u8_512 --iters 2 > /dev/null 2>&1
    echo "== $v"; python tools/launch_summary.py gpurun_out/r2s14_$v.csv 0 2>/dev/null | head -3
done
timeout -k 10 300 python bench.py --workload groupcomm_u8_512 --no-other-configs 2>/dev/null | tail -1 > gpurun_out/r2s14_bench_gc.json
python -c "
import json; d=json.load(open('gpurun_out/r2s14_bench_gc.json')); print('groupcomm %.1f mix/s %.3f ms fwd_hbm %.3f' % (d['value'], d['ms_per_step'], d['forward_hbm']['frac']))"
