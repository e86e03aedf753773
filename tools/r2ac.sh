#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for lib in "" variants/pyr5.so variants/pyr6.so; do
  echo "=== ${lib:-product (4 CTAs/SM)}"
  export SDR_B200_LIB=${lib:+$PWD/$lib}
  [ -z "$lib" ] && unset SDR_B200_LIB
  timeout -k 10 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -p no:cacheprovider -k "pyramid" 2>&1 | tail -1
  for wl in improved_u16_512 improved_u36_2048 groupcomm_u8_512; do
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-other-configs --workload $wl 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); ks=d['roofline']['kernels']
print('bench $wl %.1f mix/s %.3f ms  per_block %.3f  pyramid %.1f us merge %.1f us' % (d['value'], d['ms_per_step'], d['roofline']['per_block']['frac'], ks[1]['avg_launch_ms']*1e3, ks[2]['avg_launch_ms']*1e3))"
  done
done
} 2>&1 | tee gpurun_out/r2ac.txt
