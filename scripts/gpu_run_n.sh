#!/bin/bash
# round 2, run N: row-quad interleaved Y (16-byte loads): parity suite + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/r2n_pytest.log | tail -20
for i in 1 2; do
python bench.py --steps 30 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('kernel_ms=%.4f step_ms=%.4f it/s=%.1f e2e=%.1f clk=%s loss=%s' % (r['avg_launch_ms'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['final_loss']))
"
done 2>&1 | tee gpurun_out/r2n_bench.txt
bash scripts/ablate.sh 64 8 2>&1 | tee -a gpurun_out/r2n_bench.txt
