#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_prepost.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
