#!/bin/bash
mkdir -p gpurun_out
for ni in 0 1; do for mb in 8 7; do
  echo "== NOINFO=$ni SYNC=4 MINBLOCKS=$mb"
  ORGYM_NET_JIT_NOINFO=$ni ORGYM_NET_JIT_MINBLOCKS_STEP=$mb ORGYM_NET_JIT_SYNC=4 INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
done; done
ORGYM_NET_JIT_NOINFO=1 ORGYM_NET_JIT_SYNC=4 INFO=0 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"net_jit_step" -s 4 -c 1 python tools/net64_quick.py 2>&1 | grep -E "duration|inst_executed|dram"
