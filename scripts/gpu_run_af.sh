#!/bin/bash
# round 2, run AF: config-4 bench line incl. the plain-closure callback loop
mkdir -p gpurun_out
timeout 300 python bench.py --config 4 --steps 200 --warmup 5 --no-cpu > gpurun_out/r2af_bench_cfg4.json 2> gpurun_out/r2af_bench_cfg4.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2af_bench_cfg4.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e'], {k:r.get(k) for k in ('frac','avg_launch_ms')}, d['notes'], d.get('clocks'))
PY
tail -3 gpurun_out/r2af_bench_cfg4.err
