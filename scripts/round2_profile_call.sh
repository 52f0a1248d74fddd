#!/usr/bin/env bash
# ncu launch lists (gpu__time_duration.sum, --clock-control none) of (1) the c5 generation step and (2) three training steps
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2_c5_launches.csv \
  python bench.py --config c5 --steps 1 --warmup 1 --no-e2e --no-cpu --profile-window > gpurun_out/r2_c5_ncu.log 2>&1
tail -2 gpurun_out/r2_c5_ncu.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/r2_train_launches_after.csv \
  python scripts/probe_train_kernels.py --short > gpurun_out/r2_train_ncu_after.log 2>&1
tail -2 gpurun_out/r2_train_ncu_after.log
python scripts/probe_train_kernels.py > gpurun_out/r2_train_probe_after.json 2>&1; cat gpurun_out/r2_train_probe_after.json
exit 0
