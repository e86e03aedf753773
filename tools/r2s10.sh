#!/usr/bin/env bash
# merge kernel: runs per thread / threads per CTA (product = 128 threads x 4 runs)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in product mp2 mp8 mp64x4 product mp2 mp8 mp64x4; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv -k regex:"merge_pyramid" \
        --log-file gpurun_out/r2s10_m_$v.csv python tools/profile_forward.py --iters 2 > /dev/null 2>&1
    echo "== merge $v"; python tools/launch_summary.py gpurun_out/r2s10_m_$v.csv 0 2>/dev/null | head -3
done
for v in product mp8; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv -k regex:"merge_pyramid" \
        --log-file gpurun_out/r2s10_g_$v.csv python tools/profile_forward.py --workload groupcomm_u8_512 --iters 2 > /dev/null 2>&1
    echo "== merge groupcomm $v"; python tools/launch_summary.py gpurun_out/r2s10_g_$v.csv 0 2>/dev/null | head -3
done
timeout -k 10 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "pyramid or golden" 2>&1 | tail -2
