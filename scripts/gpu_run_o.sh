#!/bin/bash
# round 2, run O: three rotating half-tile Y buffers (setmaxnreg 32/112): parity + timing
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -q -x -k "grad_loss" -p no:cacheprovider 2>&1 | tail -1
timeout 240 python -m pytest tests -m gpu -q -x --timeout 100 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded 2>&1 | tail -1
for i in 1 2; do
timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('kernel_ms=%.4f step_ms=%.4f it/s=%.1f e2e=%.1f clk=%s loss=%s' % (r['avg_launch_ms'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['final_loss']))
"
done 2>&1 | tee gpurun_out/r2o_bench.txt
for ab in 8 64; do PMX_ABLATE=$ab timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('ablate=%3d kernel_ms=%.3f step_ms=%.3f' % ($ab, r['avg_launch_ms'], d['ms_per_step']))
"; done 2>&1 | tee -a gpurun_out/r2o_bench.txt
