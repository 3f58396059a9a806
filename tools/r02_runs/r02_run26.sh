#!/bin/bash
mkdir -p gpurun_out
for op in 1 0; do
  echo "== ONEPASS=$op"
  ORGYM_NET_JIT_ONEPASS=$op INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  ORGYM_NET_JIT_ONEPASS=$op INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  ORGYM_NET_JIT_ONEPASS=$op INFO=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,lts__t_sector_hit_rate.pct --clock-control none -k regex:"net_jit_step|net_obs" -s 8 -c 4 python tools/net64_quick.py 2>&1 | grep -E "net_jit_step|net_obs_kernel|duration|dram__|inst_executed|registers|hit_rate" | head -28
done
