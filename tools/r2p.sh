#!/usr/bin/env bash
# 2-GPU validation: DataParallel / second-device tests, N=2 bench (both arms) under torchrun
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  nvidia-smi -L
  timeout -k 10 300 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "two_devices or second_device or data_parallel" -rs 2>&1 | tail -6
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2p_bench_n2.json
  python -c "
import json; d=json.load(open('gpurun_out/r2p_bench_n2.json'))
print('N=2: %.1f mix/s %.3f ms e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print([ (o['workload'], round(o['value'],1), round(o['forward_hbm_frac'],3)) for o in d['other_configs']])"
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --impl reference --gpus 2 --steps 10 --warmup 1 2>&1 | tail -1 | cut -c1-600
} 2>&1 | tee gpurun_out/r2p.txt
