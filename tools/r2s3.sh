#!/usr/bin/env bash
# ncu --set full of the kernels furthest below their roofline: GroupComm small-channel 1x1 convs + TAC, the causal
# depthwise stage, the original model's softmax gate
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"pw_small_kernel|tac_mma16" -s 6 -c 3 \
    -o gpurun_out/prof_r2s3_gc -f python tools/profile_forward.py --workload groupcomm_u8_512 --iters 1 > gpurun_out/r2s3_gc.log 2>&1
tail -1 gpurun_out/r2s3_gc.log
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"causal_pyramid" -s 2 -c 1 \
    -o gpurun_out/prof_r2s3_causal -f python tools/profile_forward.py --workload causal_u16_512 --iters 1 > gpurun_out/r2s3_causal.log 2>&1
tail -1 gpurun_out/r2s3_causal.log
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"softmax_gate|encoder_kernel" -c 2 \
    -o gpurun_out/prof_r2s3_orig -f python tools/profile_forward.py --workload original_u16_512 --iters 1 > gpurun_out/r2s3_orig.log 2>&1
tail -1 gpurun_out/r2s3_orig.log
ls -la gpurun_out/*.ncu-rep
