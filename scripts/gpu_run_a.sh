#!/bin/bash
# round 2, run A: full GPU test suite + bench lines of every config + reference arm + sanitizer on the smoke test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2a_pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r2a_bench_n1_100.json 2>> gpurun_out/r2a_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
timeout 300 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err
timeout 300 python bench.py --config 4 --steps 200 --warmup 5 --no-cpu > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err
timeout 300 python bench.py --config 5 --K 64 --steps 10 --warmup 2 --no-cpu > gpurun_out/r2a_bench_cfg5_k64.json 2> gpurun_out/r2a_bench_cfg5.err
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_racecheck.log
tail -5 gpurun_out/r2a_pytest.log; tail -15 gpurun_out/r2a_pytest_all.log; cat gpurun_out/r2a_bench_n1.json | cut -c1-600
