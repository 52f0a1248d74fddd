#!/usr/bin/env bash
# gpurun --gpus 8 -- 'bash scripts/round2_eight_gpu.sh': the driver's scaling command at N = 8 (c3) and the c5 shape on 8 GPUs
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_zz_exchange.py -q -m gpu > gpurun_out/r2_exchange_pytest.log 2>&1; tail -2 gpurun_out/r2_exchange_pytest.log
LCREC_DIST_TIMING=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus "$N" --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "c3 rc=$?"; grep "dist timing" gpurun_out/r2_bench_${N}gpu.err | tail -2; cut -c1-400 gpurun_out/r2_bench_${N}gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --config c5 --gpus "$N" --items 250000 --steps 2 --warmup 2 --e2e-items 100000 > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
echo "c5 rc=$?"; tail -2 gpurun_out/r2_bench_c5_${N}gpu.err; cut -c1-400 gpurun_out/r2_bench_c5_${N}gpu.json
exit 0
