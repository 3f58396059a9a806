#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_net64_g.log
for cfg in "8 1" "8 2" "8 3" "7 4" "7 2" "7 3"; do set -- $cfg
  echo "== MINBLOCKS_STEP=$1 GROUP=$2" >> $L
  ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_MINBLOCKS_STEP=$1 ORGYM_NET_JIT_GROUP=$2 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
done
cat $L
