#!/usr/bin/env bash
# A/B of the tile-staged kernel's occupancy targets + the causal micro-optimisations
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "pointwise or causal or groupcomm or golden or cfg4" 2>&1 | tail -3
for v in product st36 st35 st34; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv -k regex:"pw_tile|pw_small" \
        --log-file gpurun_out/r2s6_$v.csv python tools/profile_forward.py --workload groupcomm_u8_512 --iters 2 > /dev/null 2>&1
    echo "== $v"; python tools/launch_summary.py gpurun_out/r2s6_$v.csv 0 2>/dev/null | head -4
done
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:causal_pyramid \
    --log-file gpurun_out/r2s6_cz.csv python tools/profile_forward.py --workload causal_u16_512 --iters 1 > /dev/null 2>&1
echo "== causal"; python tools/launch_summary.py gpurun_out/r2s6_cz.csv 0 2>/dev/null | head -3
for t in 160 192; do
    SDR_CZ_THREADS=$t timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:causal_pyramid \
        --log-file gpurun_out/r2s6_cz_$t.csv python tools/profile_forward.py --workload causal_u16_512 --iters 1 > /dev/null 2>&1
    echo "== causal threads $t"; python tools/launch_summary.py gpurun_out/r2s6_cz_$t.csv 0 2>/dev/null | head -3
done
