#!/bin/bash
# A/B timing of library builds on one box: scripts/ab.sh <steps> libA.so libB.so ...   (each run twice, interleaved)
steps=$1; shift
for rep in 1 2; do
for lib in "$@"; do
  cp "$lib" proxmin_b200/libproxmin_b200.so
  python bench.py --steps $steps --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('%-28s kernel_ms=%.4f step_ms=%.4f it/s=%.1f e2e=%.1f clk=%s' % ('$lib', r['avg_launch_ms'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))
"
done
done
