#!/bin/bash
# round 2, run AH: refill of the rotating Y buffers right after the conversion half that consumed them
mkdir -p gpurun_out
cp variants/lib_early.so proxmin_b200/libproxmin_b200.so
timeout 100 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or grad_loss or pgm_matches" -p no:cacheprovider --timeout 60 2>&1 | tail -1
timeout 400 bash scripts/ab.sh 30 variants/lib_late.so variants/lib_early.so 2>&1 | tee gpurun_out/r2ah_refill.txt
