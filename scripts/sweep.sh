#!/bin/bash
# usage: sweep.sh VAR v1 v2 ...   -> kernel ms for each value of env VAR
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('$var=$v kernel_ms=%.3f step_ms=%.3f frac=%.3f' % (r['avg_launch_ms'], d['ms_per_step'], r['frac']))
"
done
