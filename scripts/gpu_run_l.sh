#!/bin/bash
# round 2, run L: TMA L2 prefetch of Y with an evict_first hint, sweep of the distance (tiles ahead)
mkdir -p gpurun_out
PMX_Y_PREFETCH=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or pgm_matches" -p no:cacheprovider 2>&1 | tail -1
for pf in 0 1 2 3 4 6 0 2; do
  PMX_Y_PREFETCH=$pf python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('prefetch=$pf kernel_ms=%.4f step_ms=%.4f it/s=%.1f clk=%s' % (r['avg_launch_ms'], d['ms_per_step'], d['value'], d['clocks']['sm_mhz']))
"
done 2>&1 | tee gpurun_out/r2l_prefetch.txt
