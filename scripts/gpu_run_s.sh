#!/bin/bash
# round 2, run S: f-row tests again + ADMM pass with L2 residency policies (config 4)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -p no:cacheprovider -k "max_entropy or barzilai or dense_L or matrix_adapter or weighted or admm or sdmm" > gpurun_out/r2s_new.log 2>&1; echo "new rc=$?" >> gpurun_out/r2s_new.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2s_new.log | tail -10
for i in 1 2; do
timeout 200 python bench.py --config 4 --steps 200 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d.get('roofline') or {}; print('cfg4 it/s=%.1f ms=%.4f e2e=%.1f roofline=%s launches=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], {k:r.get(k) for k in ('frac','avg_launch_ms','achieved')}, d.get('gpu_launches')))
"
done 2>&1 | tee gpurun_out/r2s_cfg4.txt
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2s_pytest.log | tail -10
PMX_TAIL_TRACE=1 timeout 120 python bench.py --N 8192 --steps 100 --warmup 5 --no-cpu > gpurun_out/r2s_trace_n8192.log 2>&1
grep TAIL gpurun_out/r2s_trace_n8192.log | tail -8
PMX_TAIL_TRACE=1 timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r2s_trace_n65536.log 2>&1
grep TAIL gpurun_out/r2s_trace_n65536.log | tail -8
for a in "--steps 20" "--steps 200" "--N 8192 --steps 200"; do
timeout 200 python bench.py $a --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d.get('roofline') or {}; print('$a it/s=%.1f ms=%.4f e2e=%.1f kernel_ms=%.4f frac=%.3f share=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['avg_launch_ms'], r['frac'], r['kernel_share_of_step']))
"
done 2>&1 | tee gpurun_out/r2s_bench.txt
