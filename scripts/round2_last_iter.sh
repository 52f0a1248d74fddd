#!/usr/bin/env bash
# Last iteration of the round: loop parity + Sinkhorn group tests, then one bench.
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_loop_ledger.py tests/test_gpu_parity.py -x -q -m gpu -k "loop or c1 or sinkhorn_group or generate_indices or segment or full_size" > gpurun_out/last_pytest.log 2>&1
tail -3 gpurun_out/last_pytest.log
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu --no-torch-cuda --no-e2e > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
tail -c 1200 gpurun_out/last_bench.json | tr ',' '\n' | grep -E "\"2[0-6]\"|rounds|sinkhorn_rows|n_unique" | tr '\n' ' '; echo
grep -o '"value": [0-9.]*' gpurun_out/last_bench.json | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/last_bench.json | head -1
exit 0
