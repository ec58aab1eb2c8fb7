#!/bin/bash
# round 2, run AK: ticket-based tile distribution in the S role of the fused tail
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q --timeout 60 -p no:cacheprovider -k "pgm or full_size or cfg2 or cached" 2>&1 | tail -1
for a in "--steps 20" "--N 8192 --steps 200"; do
timeout 100 python bench.py $a --warmup 5 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d.get('roofline') or {}; print('$a it/s=%.1f ms=%.4f kernel_ms=%.4f share=%.3f loss=%s' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['kernel_share_of_step'], d['final_loss']))
"
done 2>&1 | tee gpurun_out/r2ak_bench.txt
PMX_TAIL_TRACE=1 timeout 60 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | grep TAIL | tail -5
