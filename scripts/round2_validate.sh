#!/usr/bin/env bash
# What the driver runs at round end, in one call: GPU test suite, smoke(), both bench arms (default config), plus c2 / c5.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_full_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_full_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r2_smoke.log | cut -c1-400
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_bench_ref.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_default.err; cut -c1-400 gpurun_out/r2_bench_default.json
timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; echo "c5 rc=$?"; cut -c1-300 gpurun_out/r2_bench_c5.json
timeout 600 python bench.py --config c2 --steps 3 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/r2_bench_c2.json
timeout 600 python bench.py --config c2 --bn --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_c2_bn.json 2> gpurun_out/r2_bench_c2_bn.err; echo "c2 bn rc=$?"; cut -c1-300 gpurun_out/r2_bench_c2_bn.json
exit 0
