#!/usr/bin/env bash
# BASELINE configs[3] at its full size: 10 M items over 8 B200 (1.25 M per GPU), the driver's torchrun command line.
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
LCREC_DIST_TIMING=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus "$N" --items 1250000 --steps 3 --warmup 3 --no-cpu --no-torch-cuda --no-e2e > gpurun_out/r2_bench_c4_10m_${N}gpu.json 2> gpurun_out/r2_bench_c4_10m_${N}gpu.err
echo "c4 rc=$?"; grep "dist timing" gpurun_out/r2_bench_c4_10m_${N}gpu.err | tail -2; cut -c1-300 gpurun_out/r2_bench_c4_10m_${N}gpu.json
exit 0
