#!/bin/bash
# round 2, run C: fused PGM tail on one GPU (parity suite + bench + old path for comparison + per-rank problem of the 8-GPU case)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_pytest.log | tail -20
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/r2c_bench_n1_200.json 2>> gpurun_out/r2c_bench_n1.err
PMX_NO_FUSED_TAIL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_bench_n1_oldtail.json 2>> gpurun_out/r2c_bench_n1.err
timeout 300 python bench.py --N 8192 --steps 200 --warmup 5 --no-cpu > gpurun_out/r2c_bench_n8192.json 2>> gpurun_out/r2c_bench_n1.err
PMX_NO_FUSED_TAIL=1 timeout 300 python bench.py --N 8192 --steps 200 --warmup 5 --no-cpu > gpurun_out/r2c_bench_n8192_oldtail.json 2>> gpurun_out/r2c_bench_n1.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 2 --no-cpu > gpurun_out/r2c_bench_cfg5.json 2> gpurun_out/r2c_bench_cfg5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_cfg5_launches.csv python bench.py --config 5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c_ncu_cfg5.log 2>&1
for f in gpurun_out/r2c_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')}, (r.get('step') or {}).get('frac'))
"; done
tail -3 gpurun_out/r2c_bench_n1.err
