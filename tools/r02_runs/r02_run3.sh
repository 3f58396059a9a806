#!/bin/bash
# round 2, GPU call 3: obs kernel v2 tests + timing; ncu full captures: net64 step pair, newsvendor rollout, serial rollout
mkdir -p gpurun_out
python -m pytest tests/test_netinv_gpu.py tests/test_canary_gpu.py tests/test_newsvendor_gpu.py -m gpu -x -q > gpurun_out/r02_tests3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests3.log
tail -4 gpurun_out/r02_tests3.log
L=gpurun_out/r02_net64_v2.log
for info in 0 1; do echo "== INFO=$info obs v2" >> $L; INFO=$info python tools/net64_quick.py >> $L 2>&1; done
echo "== INFO=0 obs v2 THREADS=256" >> $L; ORGYM_NET_JIT_THREADS=256 INFO=0 python tools/net64_quick.py >> $L 2>&1
cat $L
python tools/bench_quick.py nv 2>&1 | tee gpurun_out/r02_nv_quick.log
python tools/prof_net64.py > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'net_jit_step|net_obs_kernel' -s 10 -c 2 -o gpurun_out/r02_net64 \
      python tools/prof_net64.py > gpurun_out/ncu_net64.log 2>&1
python tools/prof_quick.py nv > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'nv_rollout_kernel' -s 0 -c 1 -o gpurun_out/r02_nv \
      python tools/prof_quick.py nv > gpurun_out/ncu_nv.log 2>&1
python tools/prof_quick.py inv > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'inv_jit_rollout' -s 1 -c 1 -o gpurun_out/r02_inv \
      python tools/prof_quick.py inv > gpurun_out/ncu_inv.log 2>&1
ls -la gpurun_out | tail -8
