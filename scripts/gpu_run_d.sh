#!/bin/bash
# round 2, run D: phase timeline of the fused tail kernel + launch lists
mkdir -p gpurun_out
PMX_TAIL_TRACE=1 timeout 300 python bench.py --N 8192 --steps 100 --warmup 5 --no-cpu > gpurun_out/r2d_trace_n8192.log 2>&1
PMX_TAIL_TRACE=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r2d_trace_n65536.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_n8192_launches.csv python bench.py --N 8192 --steps 10 --warmup 2 --no-cpu > gpurun_out/r2d_ncu.log 2>&1
grep TAIL gpurun_out/r2d_trace_n8192.log | tail -12; echo; grep TAIL gpurun_out/r2d_trace_n65536.log | tail -12
grep -E "k_pgm_tail|k_grad" gpurun_out/r2d_n8192_launches.csv | tail -6 | cut -d, -f5,15
