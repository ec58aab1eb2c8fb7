#!/bin/bash
# round 2, run AD: final verification on one B200 (what the driver runs: pytest -m gpu, smoke, bench)
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider > $O/r2ad_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2ad_pytest.log
grep -E "passed|failed|FAILED|rc=" $O/r2ad_pytest.log | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > $O/r2ad_bench_default.json 2> $O/r2ad_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2ad_bench_reference.json 2>> $O/r2ad_bench.err
for f in $O/r2ad_bench_default.json $O/r2ad_bench_reference.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('impl','value','ms_per_step','steps','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step','traffic')}, d.get('cpu_baseline'), d.get('clocks'))
PY
done
