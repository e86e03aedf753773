#!/usr/bin/env bash
# round-2 GPU call A: full GPU suite on the product, A/B of the compile-ready variants, encoder ncu capture
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r2a_smi.txt
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -rs 2>&1 | tail -40 > gpurun_out/r2a_pytest.txt
tail -3 gpurun_out/r2a_pytest.txt
timeout -k 10 1500 tools/ab_variants.sh > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:pw_mma_kernel -s 36 -c 1 \
    -o gpurun_out/prof_r2a_encoder python tools/profile_forward.py --iters 2 > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log
cat gpurun_out/ab_variants.txt
