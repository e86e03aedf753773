#!/usr/bin/env bash
# validate + measure: tile-staged small-channel kernel (GroupComm), templated softmax gate + tensor-core encoder
# (original model), rebalanced causal depthwise stage
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "pointwise or causal or original or groupcomm or golden or softmax or smoke or cfg4 or encoder" 2>&1 | tail -8
for wl in groupcomm_u8_512 causal_u16_512 original_u16_512; do
    timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
        --log-file gpurun_out/r2s4_launches_$wl.csv python tools/profile_forward.py --workload $wl --iters 2 > /dev/null 2>&1
    echo "== $wl"; python tools/launch_summary.py gpurun_out/r2s4_launches_$wl.csv | head -9
done
for t in 128 160 192 256; do
    SDR_CZ_THREADS=$t timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:causal_pyramid \
        --log-file gpurun_out/r2s4_cz_$t.csv python tools/profile_forward.py --workload causal_u16_512 --iters 1 > /dev/null 2>&1
    echo "== causal threads $t"; python tools/launch_summary.py gpurun_out/r2s4_cz_$t.csv 0 | head -3
done
for wl in groupcomm_u8_512 causal_u16_512 original_u16_512; do
    timeout -k 10 300 python bench.py --workload $wl --no-other-configs 2>/dev/null | tail -1 > gpurun_out/r2s4_bench_$wl.json
    python -c "
import json; d=json.load(open('gpurun_out/r2s4_bench_$wl.json')); print('$wl %.1f mix/s %.3f ms fwd_hbm %.3f' % (d['value'], d['ms_per_step'], d['forward_hbm']['frac']))"
done
