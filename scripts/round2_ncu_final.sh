#!/usr/bin/env bash
# Final ncu evidence of round 2 (one GPU): (1) launch list of one timed step of the default bench (after the plain run exited 0),
# (2) `ncu --set full` of the first collision round's per-group Sinkhorn kernels (3 warp kernels, 2 column kernels, literal re-run).
set -u
mkdir -p gpurun_out
python bench.py --items 262144 --steps 1 --warmup 1 --no-cpu --no-e2e --no-torch-cuda > gpurun_out/r2_final_plain.json 2> gpurun_out/r2_final_plain.err
echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2_final_launches.csv \
  python bench.py --items 262144 --steps 1 --warmup 1 --no-cpu --no-e2e --no-torch-cuda --profile-window > gpurun_out/r2_final_launches.log 2>&1
grep -c "gpu__time_duration" gpurun_out/r2_final_launches.csv
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sinkhorn_groups" -c 6 \
  -o gpurun_out/r2_sk256_v2_full python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-torch-cuda --profile-window > gpurun_out/r2_sk256_v2_full.log 2>&1
ncu -i gpurun_out/r2_sk256_v2_full.ncu-rep --page raw --csv > gpurun_out/r2_sk256_v2_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_sk256_v2_full.ncu-rep --page source --csv --kernel-name regex:"warp_kernel<2" > gpurun_out/r2_sk256_v2_warp2_source.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/r2_sk256_v2_full.ncu-rep 2>/dev/null || echo 0)
echo "report bytes: $sz"
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/r2_sk256_v2_full.ncu-rep; fi
ls -la gpurun_out/r2_sk256_v2* gpurun_out/r2_final*
exit 0
