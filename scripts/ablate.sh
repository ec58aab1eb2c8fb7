#!/bin/bash
# timing ablation of the fused gradient kernel (results are numerically meaningless, only kernel ms matters)
for ab in "$@"; do
  PMX_ABLATE=$ab python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('ablate=%3d kernel_ms=%.3f step_ms=%.3f' % ($ab, r['avg_launch_ms'], d['ms_per_step']))
"
done
