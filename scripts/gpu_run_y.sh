#!/bin/bash
# round 2, run Y (2 GPUs): sharded adaprox / bsdmm with the exchanges over peer memory
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > $O/r2y_mgpu_check.log 2>&1; echo "mgpu rc=$?" >> $O/r2y_mgpu_check.log
grep -E "ok$|FAIL|rc=|rror" $O/r2y_mgpu_check.log | tail -12
for c in 3 5; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$c bench.py --gpus 2 --config $c --steps 20 --warmup 3 --no-cpu > $O/r2y_cfg${c}_n2.json 2> $O/r2y_cfg${c}_n2.err; echo "cfg$c rc=$?"
python - $O/r2y_cfg${c}_n2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d.get('roofline') or {}
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss','replica_diff','exchange','sub_iterations')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
except Exception as e:
    print('unreadable', e)
PY
done
