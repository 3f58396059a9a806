#!/bin/bash
mkdir -p gpurun_out
for v in default t512 t512b16 b24 b40 t384; do
  echo "== variant $v"
  if [ $v = default ]; then unset ORGYM_B200_LIB; else export ORGYM_B200_LIB=/root/repo/or-gym-inventory_b200/csrc/variants/$v.so; fi
  INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  INFO=0 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:"net_obs" -s 4 -c 1 python tools/net64_quick.py 2>&1 | grep -E "duration|registers"
done
