#!/bin/bash
# round 2, run K: timing ablations of the gradient kernel (PMX_ABLATE bits: 1 no MMA1, 2 no MMA2, 4 no MMA3, 8 no Y, 16 no R^T store, 32 no flush)
mkdir -p gpurun_out
cp variants/lib_base.so proxmin_b200/libproxmin_b200.so
bash scripts/ablate.sh 0 7 39 23 8 32 16 2 4 6 15 47 2>&1 | tee gpurun_out/r2k_ablate.txt
