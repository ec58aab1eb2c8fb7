#!/bin/bash
mkdir -p gpurun_out
PMX_TAIL_TRACE=1 timeout 300 python bench.py --N 8192 --steps 100 --warmup 5 --no-cpu > gpurun_out/r2f_trace_n8192.log 2>&1
grep TAIL gpurun_out/r2f_trace_n8192.log | tail -13
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pgm_tail -s 6 -c 1 -o gpurun_out/r2f_tail python bench.py --N 8192 --steps 10 --warmup 2 --no-cpu > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
