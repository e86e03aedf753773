#!/usr/bin/env bash
# tile-staged kernel: trimmed epilogue / no activation code in the pre-add instantiation (product) vs before (variants/pwprev.so)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "pointwise or groupcomm or golden or cfg4" 2>&1 | tail -2
for v in product pwprev product pwprev; do
    LIB=""; [ "$v" != product ] && LIB="$PWD/variants/$v.so"
    SDR_B200_LIB=$LIB timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv -k regex:"pw_tile" \
        --log-file gpurun_out/r2s13_$v.csv python tools/profile_forward.py --workload groupcomm_u8_512 --iters 2 > /dev/null 2>&1
    echo "== $v"; python tools/launch_summary.py gpurun_out/r2s13_$v.csv 0 2>/dev/null | head -4
done
