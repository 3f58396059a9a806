#!/bin/bash
mkdir -p gpurun_out
for th in 128 256 512; do for sy in 0 1 4; do
  echo "== THREADS=$th SYNC=$sy"
  ORGYM_NET_JIT_THREADS=$th ORGYM_NET_JIT_SYNC=$sy INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
done; done
