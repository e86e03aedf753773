#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"dw_pyramid|merge_pyramid" -s 4 -c 2 \
    -o gpurun_out/prof_r2t_pyr -f python tools/profile_forward.py --iters 1 > gpurun_out/r2t_ncu.log 2>&1
tail -2 gpurun_out/r2t_ncu.log
