#!/bin/bash
# round 2, run V: final single-GPU measurements: parity suite, bench lines of configs 2-5, ncu captures, sanitizer
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2v_pytest.log
grep -E "passed|failed|FAILED|rc=" $O/r2v_pytest.log | tail -6
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2v_bench_cfg2.json 2> $O/r2v_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2v_bench_reference.json 2>> $O/r2v_bench.err
timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu > $O/r2v_bench_cfg2_200.json 2>> $O/r2v_bench.err
timeout 200 python bench.py --N 8192 --steps 200 --warmup 5 --no-cpu > $O/r2v_bench_cfg2_n8192.json 2>> $O/r2v_bench.err
timeout 200 python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O/r2v_bench_cfg3.json 2>> $O/r2v_bench.err
timeout 200 python bench.py --config 4 --steps 200 --warmup 5 --no-cpu > $O/r2v_bench_cfg4.json 2>> $O/r2v_bench.err
timeout 200 python bench.py --config 5 --steps 10 --warmup 2 --no-cpu > $O/r2v_bench_cfg5.json 2>> $O/r2v_bench.err
for f in $O/r2v_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], 'unreadable', e); raise SystemExit
r=d.get('roofline') or {}
print(sys.argv[1].split('/')[-1], {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), {k:r.get(k) for k in ('frac','avg_launch_ms','kernel_share_of_step')}, (r.get('step') or {}).get('frac'), d.get('cpu_baseline') and d['cpu_baseline'].get('value'))
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_grad_umma -s 6 -c 1 -o $O/r2v_grad_umma python bench.py --steps 6 --warmup 2 --no-cpu > $O/r2v_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pgm_tail -s 6 -c 1 -o $O/r2v_pgm_tail python bench.py --steps 6 --warmup 2 --no-cpu > $O/r2v_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_admm_pass -s 6 -c 1 -o $O/r2v_admm_pass python bench.py --config 4 --steps 12 --warmup 3 --no-cpu > $O/r2v_ncu3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2v_launches.csv python bench.py --steps 10 --warmup 2 --no-cpu > $O/r2v_ncu4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2v_n8192_launches.csv python bench.py --N 8192 --steps 10 --warmup 2 --no-cpu > $O/r2v_ncu5.log 2>&1
ls -la $O/*.ncu-rep
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_sanitizer_memcheck.txt 2>&1; tail -3 $O/r2v_sanitizer_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_sanitizer_racecheck.txt 2>&1; tail -3 $O/r2v_sanitizer_racecheck.txt
