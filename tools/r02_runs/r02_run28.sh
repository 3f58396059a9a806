#!/bin/bash
mkdir -p gpurun_out
for mb in 8 7 6; do for ah in 0 1 2; do
  echo "== MINBLOCKS=$mb AHEAD=$ah"
  ORGYM_NET_JIT_MINBLOCKS_STEP=$mb ORGYM_NET_JIT_AHEAD=$ah INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
done; done
