#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_invmgmt_gpu.py -m gpu -x -q -k "specialised" > gpurun_out/r02_tests15.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests15.log
tail -3 gpurun_out/r02_tests15.log
for cfg in "0 4" "4 4" "4 5" "8 4" "8 5" "12 4"; do set -- $cfg
  echo "== RND_UNROLL=$1 MINBLOCKS=$2"
  ORGYM_JIT_CACHE=0 ORGYM_INV_JIT_RND_UNROLL=$1 ORGYM_INV_JIT_MINBLOCKS=$2 python tools/bench_quick.py inv 2>&1 | grep "rollout random"
done
