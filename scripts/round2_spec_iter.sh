#!/usr/bin/env bash
# Speculative later rounds (lcrec_indexer_set_speculative): parity tests + A/B of the bench stages 21 / 23.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loop_ledger.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/spec_pytest.log 2>&1
tail -3 gpurun_out/spec_pytest.log
i=0
for s in 2 0 2; do
  i=$((i+1))
  LCREC_SPECULATIVE=$s python bench.py --steps 4 --warmup 3 --no-cpu --no-torch-cuda --no-e2e > gpurun_out/spec_bench_${s}_$i.json 2> gpurun_out/spec_bench_${s}_$i.err
  tail -c 1200 gpurun_out/spec_bench_${s}_$i.json | tr ',' '\n' | grep -E "\"2[0-6]\"|rounds|sinkhorn_rows|n_unique" | tr '\n' ' '; echo
  grep -o '"value": [0-9.]*' gpurun_out/spec_bench_${s}_$i.json | head -1
done
exit 0
