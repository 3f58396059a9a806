#!/bin/bash
mkdir -p gpurun_out
for v in t384b42 t384b14 t448b18 t320b26 t384b22 t384b21d6; do
  echo "== variant $v"
  export ORGYM_B200_LIB=/root/repo/or-gym-inventory_b200/csrc/variants/$v.so
  INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  INFO=0 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:"net_obs" -s 4 -c 1 python tools/net64_quick.py 2>&1 | grep -E "duration|registers"
done
