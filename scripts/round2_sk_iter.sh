#!/usr/bin/env bash
# Iteration loop for the 256-code per-group Sinkhorn kernels: parity tests, a short bench, the launch list of one step.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_loop_ledger.py -x -q -m gpu > gpurun_out/sk_iter_pytest.log 2>&1
tail -3 gpurun_out/sk_iter_pytest.log
python bench.py --steps 3 --warmup 2 --no-cpu --no-torch-cuda > gpurun_out/sk_iter_bench.json 2> gpurun_out/sk_iter_bench.err
tail -c 1500 gpurun_out/sk_iter_bench.json | tr ',' '\n' | grep -E "stage_ms|\"2[0-9]\"|value|ms_per_step" | head -20
ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio \
  --clock-control none --profile-from-start off -k regex:"sinkhorn_groups" -c 60 --csv --log-file gpurun_out/sk_iter_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-torch-cuda --profile-window > gpurun_out/sk_iter_ncu.log 2>&1
grep -c sinkhorn gpurun_out/sk_iter_launches.csv
exit 0
