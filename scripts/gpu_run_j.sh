#!/bin/bash
# round 2, run J: A/B of gradient-kernel variants (Y loads issued before the R^T wait; stacked [A_hi|A_lo] G_S GEMM)
mkdir -p gpurun_out
for v in base yearly stack both; do
  cp variants/lib_$v.so proxmin_b200/libproxmin_b200.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or pgm_matches" -p no:cacheprovider 2>&1 | tail -1 | sed "s/^/$v: /"
done
bash scripts/ab.sh 30 variants/lib_base.so variants/lib_yearly.so variants/lib_stack.so variants/lib_both.so 2>&1 | tee gpurun_out/r2j_ab.txt
