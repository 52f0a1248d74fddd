#!/usr/bin/env bash
# Multi-GPU checks of round 2 in ONE call:   gpurun --gpus 2 --timeout 900 -- 'bash scripts/round2_two_gpu.sh'
#   1. the two pytest cases that need 2 GPUs (sharded generation == single GPU bitwise, row-sharded Sinkhorn == single GPU)
#   2. the row-sharded Sinkhorn back to back without a barrier (the absolute-step slot parity fixed in round 2)
#   3. the data-parallel trainer: toy trajectory vs the reference + the run.sh shape vs the reference's per-step losses, bn False
#      and True (synchronised BatchNorm), replicas bit-identical, step time
#   4. bench.py on N GPUs (short): value, e2e, h2d probe, "sharded_equals_single"
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
echo "GPUs: $N"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sharded_generation_equals_single_gpu or row_sharded_sinkhorn" > gpurun_out/r2_2gpu_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_2gpu_pytest.log; tail -3 gpurun_out/r2_2gpu_pytest.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29521 \
  scripts/check_dist_sinkhorn.py > gpurun_out/r2_dist_sinkhorn_${N}gpu.json 2> gpurun_out/r2_dist_sinkhorn.err
echo "dist sinkhorn rc=$?"; cat gpurun_out/r2_dist_sinkhorn_${N}gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29522 \
  scripts/check_dp_trainer.py > gpurun_out/r2_dp_trainer_${N}gpu.json 2> gpurun_out/r2_dp_trainer.err
echo "dp trainer rc=$?"; tail -3 gpurun_out/r2_dp_trainer.err; cat gpurun_out/r2_dp_trainer_${N}gpu.json
LCREC_DIST_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29523 \
  bench.py --gpus "$N" --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?"; grep "dist timing" gpurun_out/r2_bench_${N}gpu.err | tail -2; cat gpurun_out/r2_bench_${N}gpu.json
exit 0
