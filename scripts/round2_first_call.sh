#!/usr/bin/env bash
# Everything that was written after round 1's GPU budget ran out, in ONE call (1 GPU part, then the 2-GPU part if visible):
#   gpurun --timeout 300 -- 'bash scripts/round2_first_call.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 300 -- 'bash scripts/round2_first_call.sh'   (adds the data-parallel trainer check)
# Outputs land in gpurun_out/r2_first_*.log / .jsonl.
set -u
mkdir -p gpurun_out
# 1. device k-means++ seeding (lcrec_kmeanspp_seed, backend device_seed) against scikit-learn
timeout 90 python -m pytest tests/pending_gpu/kmeanspp_seed_check.py -q -m gpu > gpurun_out/r2_first_kmeanspp.log 2>&1
echo "rc=$?" >> gpurun_out/r2_first_kmeanspp.log
# 2. one small call of every kernel added late in round 1 (odd sizes, empty tensors, flushes, relocation)
timeout 60 python scripts/sanitize_new_kernels.py > gpurun_out/r2_first_smallcalls.log 2>&1
echo "rc=$?" >> gpurun_out/r2_first_smallcalls.log
# 3. data-parallel trainer against the reference loss trajectory + step time, when 2+ GPUs are visible
NGPU=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$NGPU" -ge 2 ]; then
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/check_dp_trainer.py > gpurun_out/r2_first_dp_trainer.jsonl 2> gpurun_out/r2_first_dp_trainer.err
  echo "rc=$?" >> gpurun_out/r2_first_dp_trainer.err
fi
tail -3 gpurun_out/r2_first_kmeanspp.log gpurun_out/r2_first_smallcalls.log
[ -f gpurun_out/r2_first_dp_trainer.jsonl ] && cat gpurun_out/r2_first_dp_trainer.jsonl
exit 0
