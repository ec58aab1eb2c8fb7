#!/bin/bash
# scripts/ab_env.sh <steps> "<ENV=val ...>" lib.so ... : kernel ms for each library under the given environment
steps=$1; shift
envs=$1; shift
for lib in "$@"; do
  cp "$lib" proxmin_b200/libproxmin_b200.so
  env $envs python bench.py --steps $steps --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('%-22s %-20s kernel_ms=%.4f step_ms=%.4f it/s=%.1f' % ('$lib', '$envs', r['avg_launch_ms'], d['ms_per_step'], d['value']))
"
done
