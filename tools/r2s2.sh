#!/usr/bin/env bash
# launch lists (ncu, one metric) of the configs that sit furthest below their roofline model
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "residual_norm" 2>&1 | tail -3
for wl in groupcomm_u8_512 causal_u16_512 original_u16_512 improved_u36_2048; do
    B=0; [ "$wl" = improved_u36_2048 ] && B=16
    timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
        --log-file gpurun_out/r2s2_launches_$wl.csv python tools/profile_forward.py --workload $wl --iters 2 --batch $B > /dev/null 2>&1
    echo "== $wl"; python tools/launch_summary.py gpurun_out/r2s2_launches_$wl.csv | head -16
done
