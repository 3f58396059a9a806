#!/bin/bash
set -u
mkdir -p gpurun_out
cap() {  # name, prof_r02 mode, kernel regex, skip, count
  python tools/prof_r02.py $2 > /dev/null 2>&1 || { echo "prof $2 failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o gpurun_out/r02_$1 \
      python tools/prof_r02.py $2 > gpurun_out/ncu_$1.log 2>&1
  echo "captured $1"
}
cap inv_lost    inv_lost    'inv_jit_rollout_bs'  1 1
cap inv_backlog inv_backlog 'inv_jit_rollout_bs'  1 1
cap inv_random  inv_random  'inv_jit_rollout_rnd' 1 1
cap inv_wide    inv_wide    'inv_rollout_kernel'  1 1
