#!/bin/bash
# round 2, run W: L2 residency policies of the ADMM pass (config 4)
mkdir -p gpurun_out
for m in 0 1 2 3 1 0; do
PMX_ADMM_L2=$m timeout 200 python bench.py --config 4 --steps 200 --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d.get('roofline') or {}; print('l2_mode=$m it/s=%.1f ms=%.4f pass_ms=%.4f' % (d['value'], d['ms_per_step'], r.get('avg_launch_ms')))
"
done 2>&1 | tee gpurun_out/r2w_admm_l2.txt
