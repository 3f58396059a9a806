#!/bin/bash
mkdir -p gpurun_out
for pp in 1 0; do
  echo "== PIPE=$pp"
  ORGYM_NET_JIT_PIPE=$pp INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  ORGYM_NET_JIT_PIPE=$pp INFO=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"net_jit_step" -s 4 -c 2 python tools/net64_quick.py 2>&1 | grep -E "duration|dram__|inst_executed|registers|shared_mem|warps_active" | head -14
done
