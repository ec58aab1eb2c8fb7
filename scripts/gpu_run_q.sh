#!/bin/bash
# round 2, run Q: test_wait spin loops for the issuer warps (v2), the accumulator wait of the residual warps (v3), both (v4)
mkdir -p gpurun_out
for v in v1 v2 v3 v4; do
  cp variants/lib_$v.so proxmin_b200/libproxmin_b200.so
  timeout 60 python -m pytest tests/test_gpu_parity.py -q -x -k "grad_loss or tcgen05" -p no:cacheprovider 2>&1 | tail -1 | sed "s/^/$v: /"
done
for rep in 1 2; do for v in v1 v2 v3 v4; do
  cp variants/lib_$v.so proxmin_b200/libproxmin_b200.so
  timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('%-8s kernel_ms=%.4f step_ms=%.4f it/s=%.1f' % ('$v', r['avg_launch_ms'], d['ms_per_step'], d['value']))
"
done; done 2>&1 | tee gpurun_out/r2q_spin.txt
