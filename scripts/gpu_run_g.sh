#!/bin/bash
# round 2, run G: fused tail with k_tail_final on the side stream (next to the gradient kernel on 147 SMs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2g_pytest.log | tail -20
PMX_TAIL_TRACE=1 timeout 300 python bench.py --N 8192 --steps 100 --warmup 5 --no-cpu > gpurun_out/r2g_trace_n8192.log 2>&1
grep TAIL gpurun_out/r2g_trace_n8192.log | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/r2g_bench_n1_200.json 2>> gpurun_out/r2g_bench_n1.err
timeout 300 python bench.py --N 8192 --steps 200 --warmup 5 --no-cpu > gpurun_out/r2g_bench_n8192.json 2>> gpurun_out/r2g_bench_n1.err
for f in gpurun_out/r2g_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','final_loss')}, 'e2e', d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')}, (r.get('step') or {}).get('frac'))
"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2g_n8192_launches.csv python bench.py --N 8192 --steps 6 --warmup 2 --no-cpu > gpurun_out/r2g_ncu.log 2>&1
grep -E "k_pgm_tail|k_grad|k_tail_final" gpurun_out/r2g_n8192_launches.csv | tail -6 | cut -d, -f5,15
tail -3 gpurun_out/r2g_bench_n1.err
