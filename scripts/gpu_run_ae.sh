#!/bin/bash
# round 2, run AE (2 GPUs): final sharded sanity check on the committed code
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > $O/r2ae_mgpu_check.log 2>&1; echo "mgpu rc=$?" >> $O/r2ae_mgpu_check.log
grep -E "ok$|FAIL|rc=|rror" $O/r2ae_mgpu_check.log | tail -10
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2ae_bench_n2.json 2> $O/r2ae_bench_n2.err; echo "bench rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2ae_ref_n2.json 2> $O/r2ae_ref_n2.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2ae_bench_n2.json','gpurun_out/r2ae_ref_n2.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline') or {}
    print({k:d.get(k) for k in ('impl','value','ms_per_step','replica_diff','exchange')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms')}, d.get('clocks'), (d.get('cpu_baseline') or {}).get('cores'))
PY
