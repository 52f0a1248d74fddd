#!/usr/bin/env bash
# `ncu --set full` capture of the first collision round's per-group Sinkhorn kernels at the default shape (256 codes x 32 dims):
# the three warp kernels (<= 2, 3-4, 5-8 rows) and the CTA kernel (>= 9 rows).  One GPU.
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sinkhorn_groups" -c 4 \
  -o gpurun_out/r2_sk256_full python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --profile-window > gpurun_out/r2_sk256_full.log 2>&1
ncu -i gpurun_out/r2_sk256_full.ncu-rep --page raw --csv > gpurun_out/r2_sk256_full_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/r2_sk256_full.ncu-rep 2>/dev/null || echo 0)
echo "report bytes: $sz"
if [ "$sz" -gt 45000000 ]; then
  ncu -i gpurun_out/r2_sk256_full.ncu-rep --page source --csv > gpurun_out/r2_sk256_full_source.csv 2>/dev/null
  rm -f gpurun_out/r2_sk256_full.ncu-rep
fi
ls -la gpurun_out/r2_sk256*
exit 0
