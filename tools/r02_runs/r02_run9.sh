#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_net64_mb.log
for cfg in "128 5" "128 6" "128 8" "256 3" "256 4" "64 8" "64 12"; do set -- $cfg
  echo "== THREADS=$1 MINBLOCKS_STEP=$2" >> $L
  ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_THREADS=$1 ORGYM_NET_JIT_MINBLOCKS_STEP=$2 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
done
for g in 2 6; do echo "== GROUP=$g" >> $L; ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_GROUP=$g INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L; done
cat $L
python tools/bench_quick.py nv 2>&1 | grep rollout
