#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:pw_mma_kernel -s 38 -c 2 \
    -o gpurun_out/prof_r2e_gemm -f python tools/profile_forward.py --iters 2 > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
