#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SM=$(nvidia-smi --query-gpu=clocks.max.sm --format=csv,noheader,nounits | head -1)
SDR_B200_LIB=$PWD/variants/trace.so SM_GHZ=1.9 timeout -k 10 200 python tools/trace_gemm.py 2>&1 | tee gpurun_out/r2g_trace.txt | tail -70
