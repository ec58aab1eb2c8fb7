#!/bin/bash
# round 2, run AJ (2 GPUs): loss after a sharded fused-tail solve (plan detached from the arena), 20-step bench
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2aj_bench_n2.json 2> gpurun_out/r2aj_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aj_bench_n2.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','final_loss','replica_diff','exchange')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms')}, d.get('clocks'))
PY
tail -2 gpurun_out/r2aj_bench_n2.err
