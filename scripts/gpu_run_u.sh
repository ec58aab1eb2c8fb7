#!/bin/bash
# round 2, run U (8 GPUs): sharded parity check (2 ranks) + scaling of config 2 at 8 / 4 / 2 / 1 GPUs
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/r2u_mgpu_check.log 2>&1; echo "mgpu rc=$?" >> gpurun_out/r2u_mgpu_check.log
grep -E "ok$|FAIL|rc=" gpurun_out/r2u_mgpu_check.log | tail -12
for cfg in "8 200" "8 20" "4 200" "2 200"; do set -- $cfg; n=$1; k=$2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps $k --warmup 5 --no-cpu > gpurun_out/r2u_bench_n${n}_$k.json 2> gpurun_out/r2u_bench_n${n}_$k.err; echo "bench $n/$k rc=$?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2u_bench_n${n}_$k.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss','replica_diff','exchange')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
"
done
timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/r2u_bench_n1_200.json 2>/dev/null
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2u_bench_n1_200.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('N=1', {k:d.get(k) for k in ('value','ms_per_step','final_loss')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
"
