#!/bin/bash
# round 2, run I (8 GPUs): scaling of config 2 with the fused tail
mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 200 --warmup 5 --no-cpu > gpurun_out/r2i_bench_n$n.json 2> gpurun_out/r2i_bench_n$n.err; echo "bench $n rc=$?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2i_bench_n$n.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss','replica_diff','exchange')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
"
tail -3 gpurun_out/r2i_bench_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2i_bench_n8_20.json 2> gpurun_out/r2i_bench_n8_20.err; echo "bench 8/20 rc=$?"
tail -c 600 gpurun_out/r2i_bench_n8_20.json | head -c 300
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2i_bench_n8_20.json').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss','replica_diff','exchange')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')})
"
