#!/usr/bin/env bash
# causal depthwise stage: level parameters staged in shared memory (product) vs loaded per level (variants/czprev.so)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "causal" 2>&1 | tail -3
for v in product czprev product czprev; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv -k regex:"causal_pyramid" \
        --log-file gpurun_out/r2s8_$v.csv python tools/profile_forward.py --workload causal_u16_512 --iters 1 > /dev/null 2>&1
    echo "== $v"; python tools/launch_summary.py gpurun_out/r2s8_$v.csv 0 2>/dev/null | head -3
done
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"causal_pyramid" -s 2 -c 1 \
    -o gpurun_out/prof_r2s8_causal -f python tools/profile_forward.py --workload causal_u16_512 --iters 1 > gpurun_out/r2s8_causal.log 2>&1
