#!/bin/bash
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:nv_level -c 2 python tools/prof_r02.py nv 2>&1 | grep -v "^==PROF" | tail -30
