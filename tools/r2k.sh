#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for wl in groupcomm_u8_512 improved_u16_512; do
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r2k_launches_$wl.csv python tools/profile_forward.py --workload $wl --iters 2 > gpurun_out/r2k_$wl.log 2>&1
tail -1 gpurun_out/r2k_$wl.log
done
