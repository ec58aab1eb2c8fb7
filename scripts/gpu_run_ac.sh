#!/bin/bash
# round 2, run AC (8 GPUs): sharded parity check + configs 3 / 5 with the peer exchanges, config 2 sanity
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > $O/r2ac_mgpu_check.log 2>&1; echo "mgpu rc=$?" >> $O/r2ac_mgpu_check.log
grep -E "ok$|FAIL|rc=|rror" $O/r2ac_mgpu_check.log | tail -10
run() { # n config steps warmup tag
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2957$1 bench.py --gpus $1 --config $2 --steps $3 --warmup $4 --no-cpu > $O/r2ac_$5.json 2> $O/r2ac_$5.err; echo "$5 rc=$?"
python - $O/r2ac_$5.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline') or {}
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss','replica_diff','exchange','sub_iterations')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
except Exception as e:
    print('unreadable', e)
PY
}
run 8 3 30 3 cfg3_n8
run 4 3 30 3 cfg3_n4
run 8 5 10 2 cfg5_n8
run 8 2 20 5 cfg2_n8_20
run 8 2 200 5 cfg2_n8_200
