#!/bin/bash
# round 2, run AB: adaprox loop after vectorised sub-iteration kernel, fused split, 32-bit index math (one GPU)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -k "adaprox or amsgrad or weighted or parabola or callbacks or pgm_matches or accel or bsdmm" > $O/r2ab_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2ab_pytest.log
grep -E "passed|failed|FAILED|rc=" $O/r2ab_pytest.log | tail -8
for i in 1 2; do
timeout 200 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d.get('roofline') or {}; print('cfg3 it/s=%.1f ms=%.4f e2e=%.1f kernel_ms=%.4f share=%.3f launches=%s sub=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['avg_launch_ms'], r['kernel_share_of_step'], d['gpu_launches'], d.get('sub_iterations')))
"
done 2>&1 | tee $O/r2ab_cfg3.txt
