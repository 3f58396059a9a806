#!/bin/bash
mkdir -p gpurun_out
for fe in 0 1; do for sy in 2 4 8; do
  echo "== FENCE=$fe SYNC=$sy"
  ORGYM_NET_JIT_FENCE=$fe ORGYM_NET_JIT_SYNC=$sy INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
done; done
echo "== FENCE=1 SYNC=0"
ORGYM_NET_JIT_FENCE=1 ORGYM_NET_JIT_SYNC=0 INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
echo "== twopass SYNC=1"
ORGYM_NET_JIT_ONEPASS=0 ORGYM_NET_JIT_SYNC=1 INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
