#!/usr/bin/env bash
# `ncu --set full` captures of the round's kernels (one GPU; each capture replays its kernel ~40 times):
#   1. linear_pair_kernel, encoder layers 1 and 2 at 131 072 rows per launch (the roofline kernel of bench.py)
#   2. sinkhorn_dense_cluster_kernel (training forward, batch 1024 x 256 codes)
#   3. sinkhorn_widereg / sinkhorn_wide kernels + wide_distances_kernel (c5, round 1 of the collision loop)
# Raw pages are exported as CSV next to the reports; profiles/ keeps the CSV summaries.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --profile-from-start off"
$NCU -k regex:linear_pair -c 2 -o gpurun_out/r2_pair_full python bench.py --items 262144 --steps 1 --warmup 1 --no-cpu --no-e2e --profile-window > gpurun_out/r2_pair_full.log 2>&1
ncu -i gpurun_out/r2_pair_full.ncu-rep --page raw --csv > gpurun_out/r2_pair_full_raw.csv 2>/dev/null
$NCU -k regex:sinkhorn_dense_cluster -c 1 -o gpurun_out/r2_dense_cluster_full python scripts/probe_train_kernels.py --short > gpurun_out/r2_dense_cluster_full.log 2>&1
ncu -i gpurun_out/r2_dense_cluster_full.ncu-rep --page raw --csv > gpurun_out/r2_dense_cluster_full_raw.csv 2>/dev/null
$NCU -k regex:"sinkhorn_wide|wide_distances" -c 6 -o gpurun_out/r2_wide_full python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e --profile-window > gpurun_out/r2_wide_full.log 2>&1
ncu -i gpurun_out/r2_wide_full.ncu-rep --page raw --csv > gpurun_out/r2_wide_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2_wide_full.ncu-rep        # 55 MB with sources: the pull-back limit is 64 MiB; the raw page is what is kept
ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv
# launch list of the default bench command (the driver's), for the share-of-step check
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2000 --csv --log-file gpurun_out/r2_default_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --profile-window > gpurun_out/r2_default_ncu.log 2>&1
tail -1 gpurun_out/r2_default_ncu.log | cut -c1-200
exit 0
