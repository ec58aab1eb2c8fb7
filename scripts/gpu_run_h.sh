#!/bin/bash
# round 2, run H (2 GPUs): sharded fused tail (reduce-scatter / all-gather over peer memory) against the oracle + bench
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/r2h_mgpu_check.log 2>&1; echo "mgpu rc=$?" >> gpurun_out/r2h_mgpu_check.log
grep -E "ok$|FAIL|rc=|Error|error" gpurun_out/r2h_mgpu_check.log | tail -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2h_bench_n2.json
tail -5 gpurun_out/r2h_bench_n2.err
