#!/usr/bin/env bash
# PASS-0 chunk size sweep of the default bench (device-resident value + e2e), one box, back to back.
set -u
mkdir -p gpurun_out
for c in 131072 262144 524288 131072; do
  python bench.py --steps 4 --warmup 3 --no-cpu --no-torch-cuda --chunk-rows $c > gpurun_out/chunk_$c.json 2> gpurun_out/chunk_$c.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/chunk_$c.json').read().strip().splitlines()[-1])
st=d['roofline']['stage_ms_per_step']
print($c, round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {k:st[k] for k in ('0','1','2','3','4','5','6','7','17','20')}, d['clocks']['sm_mhz'])
PY
done
exit 0
