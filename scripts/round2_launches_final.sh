#!/usr/bin/env bash
# ncu launch list of one timed step of the default bench on the final code (after the plain run exited 0).
set -u
mkdir -p gpurun_out
python bench.py --items 262144 --steps 1 --warmup 2 --no-cpu --no-e2e --no-torch-cuda > gpurun_out/r2_final2_plain.json 2> gpurun_out/r2_final2_plain.err
echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/r2_final2_launches.csv \
  python bench.py --items 262144 --steps 1 --warmup 2 --no-cpu --no-e2e --no-torch-cuda --profile-window > gpurun_out/r2_final2_launches.log 2>&1
echo "ncu rc=$?"; grep -c "gpu__time_duration" gpurun_out/r2_final2_launches.csv
exit 0
