#!/bin/bash
# round 2, run R: SURVEY 8-f rows on the GPU (max entropy, BB stepper, dense L, weighted likelihood) + full parity suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -p no:cacheprovider -k "max_entropy or barzilai or dense_L or matrix_adapter or weighted" > gpurun_out/r2r_new.log 2>&1; echo "new rc=$?" >> gpurun_out/r2r_new.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/r2r_new.log | tail -20
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_multi_gpu_sharded > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2r_pytest.log | tail -20
