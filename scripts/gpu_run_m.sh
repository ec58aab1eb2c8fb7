#!/bin/bash
# round 2, run M: what the Y loads cost -- DRAM latency or load issue?  (64: L2-resident Y, 256: half of the loads)
mkdir -p gpurun_out
bash scripts/ablate.sh 0 64 256 320 71 263 327 2>&1 | tee gpurun_out/r2m_ablate.txt
