#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/bench_weighted.py 2>&1 | tee gpurun_out/r2ag_weighted.txt
