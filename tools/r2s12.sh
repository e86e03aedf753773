#!/usr/bin/env bash
# 2-GPU sanity: the tests that need two devices, and the scaling bench under torchrun (N = 2)
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -k 10 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k "two_devices or second_device or data_parallel" 2>&1 | tail -4
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/r2s12_bench2.err | tail -1 > gpurun_out/r2s12_bench2.json
tail -3 gpurun_out/r2s12_bench2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2s12_bench2.json'))
print('N=2: %.1f mix/s %.3f ms e2e %.1f n_gpus %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus']))
print('others', [(o['workload'], round(o['value'], 1)) for o in d['other_configs']])
PY
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
