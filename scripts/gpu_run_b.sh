#!/bin/bash
# round 2, run B: K > 64 tcgen05 path (tests + config-5 bench line) and a regression pass of the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
timeout 300 python bench.py --config 5 --steps 10 --warmup 2 --no-cpu > gpurun_out/r2b_bench_cfg5.json 2> gpurun_out/r2b_bench_cfg5.err
timeout 300 python bench.py --config 5 --K 64 --steps 10 --warmup 2 --no-cpu > gpurun_out/r2b_bench_cfg5_k64.json 2>> gpurun_out/r2b_bench_cfg5.err
grep -E "passed|failed|FAILED|Error" gpurun_out/r2b_pytest.log | tail -20
for f in gpurun_out/r2b_bench_n1.json gpurun_out/r2b_bench_cfg5.json gpurun_out/r2b_bench_cfg5_k64.json; do cut -c1-300 $f; done
tail -5 gpurun_out/r2b_bench_cfg5.err
