#!/usr/bin/env bash
# A/B: TAC compiled for 3 resident CTAs; merge kernel items / threads
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in product tac3 product tac3; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv -k regex:"tac_mma16" \
        --log-file gpurun_out/r2s9_$v.csv python tools/profile_forward.py --workload groupcomm_u8_512 --iters 2 > /dev/null 2>&1
    echo "== $v"; python tools/launch_summary.py gpurun_out/r2s9_$v.csv 0 2>/dev/null | head -3
done
for v in product mp1 mp4 mp256 product mp1 mp4 mp256; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv -k regex:"merge_pyramid" \
        --log-file gpurun_out/r2s9_m_$v.csv python tools/profile_forward.py --iters 2 > /dev/null 2>&1
    echo "== merge $v"; python tools/launch_summary.py gpurun_out/r2s9_m_$v.csv 0 2>/dev/null | head -3
done
SDR_B200_LIB=$PWD/variants/tac3.so timeout -k 10 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "tac or groupcomm" 2>&1 | tail -2
