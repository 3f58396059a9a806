#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_net64_g2.log
for cfg in "128 7 1" "128 6 1" "64 16 1" "64 14 1" "256 4 1" "128 9 1"; do set -- $cfg
  echo "== THREADS=$1 MINBLOCKS_STEP=$2 GROUP=$3" >> $L
  ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_THREADS=$1 ORGYM_NET_JIT_MINBLOCKS_STEP=$2 ORGYM_NET_JIT_GROUP=$3 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
done
echo "== 128 8 1 with PREFETCH=1" >> $L; ORGYM_NET_JIT_PREFETCH=1 ORGYM_NET_JIT_MINBLOCKS_STEP=8 ORGYM_NET_JIT_GROUP=1 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
echo "== 128 8 1 with PREFETCH=3" >> $L; ORGYM_NET_JIT_PREFETCH=3 ORGYM_NET_JIT_MINBLOCKS_STEP=8 ORGYM_NET_JIT_GROUP=1 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
grep -E "==|step" $L
