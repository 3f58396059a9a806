#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_invmgmt_gpu.py -m gpu -x -q -k "specialised" > gpurun_out/r02_tests17.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests17.log
tail -3 gpurun_out/r02_tests17.log
for u in -1 8 10 12 15; do echo "== RND_UNROLL=$u"; if [ $u -lt 0 ]; then ORGYM_JIT_CACHE=0 python tools/bench_quick.py inv 2>&1 | grep "rollout"; else ORGYM_JIT_CACHE=0 ORGYM_INV_JIT_RND_UNROLL=$u python tools/bench_quick.py inv 2>&1 | grep "rollout random"; fi; done
cap() {  # name, prof_r02 mode, kernel regex, skip, count
  python tools/prof_r02.py $2 > /dev/null 2>&1 || { echo "prof $2 failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o gpurun_out/r02_$1 \
      python tools/prof_r02.py $2 > gpurun_out/ncu_$1.log 2>&1
}
cap inv_lost    inv_lost    'inv_jit_rollout_bs'  1 1
cap inv_backlog inv_backlog 'inv_jit_rollout_bs'  1 1
cap inv_random  inv_random  'inv_jit_rollout_rnd' 1 1
cap inv_wide    inv_wide    'inv_rollout_kernel'  1 1
cap inv_step    inv_step    'inv_step_kernel'     13 1
ls -la gpurun_out/*.ncu-rep | tail -6
