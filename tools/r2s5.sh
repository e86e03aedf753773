#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"pw_tile_kernel" -s 2 -c 2 \
    -o gpurun_out/prof_r2s5_gc -f python tools/profile_forward.py --workload groupcomm_u8_512 --iters 1 > gpurun_out/r2s5_gc.log 2>&1
tail -1 gpurun_out/r2s5_gc.log
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"causal_pyramid" -s 2 -c 1 \
    -o gpurun_out/prof_r2s5_causal -f python tools/profile_forward.py --workload causal_u16_512 --iters 1 > gpurun_out/r2s5_causal.log 2>&1
tail -1 gpurun_out/r2s5_causal.log
