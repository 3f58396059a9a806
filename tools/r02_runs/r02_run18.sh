#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_netinv_gpu.py tests/test_canary_gpu.py tests/test_random_configs_gpu.py -m gpu -x -q > gpurun_out/r02_tests18.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests18.log
tail -3 gpurun_out/r02_tests18.log
L=gpurun_out/r02_net64_ldgsts.log
for cfg in "0 8 1" "1 8 1" "1 8 2" "1 7 1" "1 8 4" "1 10 1" "1 12 1"; do set -- $cfg
  echo "== LDGSTS=$1 MINBLOCKS_STEP=$2 GROUP=$3" >> $L
  ORGYM_NET_JIT_LDGSTS=$1 ORGYM_NET_JIT_MINBLOCKS_STEP=$2 ORGYM_NET_JIT_GROUP=$3 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
done
grep -E "==|step" $L
