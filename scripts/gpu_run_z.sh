#!/bin/bash
# round 2, run Z (2 GPUs): adaprox without host round trips inside an iteration (speculative sub-iterations + pause)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -k "adaprox or amsgrad or weighted or parabola or callbacks" > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log
grep -E "passed|failed|FAILED|rc=" $O/r2z_pytest.log | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > $O/r2z_mgpu_check.log 2>&1; echo "mgpu rc=$?" >> $O/r2z_mgpu_check.log
grep -E "ok$|FAIL|rc=|rror" $O/r2z_mgpu_check.log | tail -12
timeout 200 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O/r2z_cfg3_n1.json 2> $O/r2z_cfg3_n1.err; echo "cfg3 n1 rc=$?"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --config 3 --steps 30 --warmup 3 --no-cpu > $O/r2z_cfg3_n2.json 2> $O/r2z_cfg3_n2.err; echo "cfg3 n2 rc=$?"
for f in $O/r2z_cfg3_n1.json $O/r2z_cfg3_n2.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d.get('roofline') or {}
    print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss','replica_diff','exchange','sub_iterations')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
except Exception as e:
    print('unreadable', e)
PY
done
