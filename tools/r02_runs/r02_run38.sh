#!/bin/bash
# final captures for the kernels whose sources changed after the first final pass (newsvendor, network)
set -u
mkdir -p gpurun_out
cap() {  # name, prof_r02 mode, kernel regex, skip, count
  python tools/prof_r02.py $2 > /dev/null 2>&1 || { echo "prof $2 failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o gpurun_out/r02_$1 \
      python tools/prof_r02.py $2 > gpurun_out/ncu_$1.log 2>&1
  echo "captured $1"
}
cap nv          nv          'nv_level_kernel|nv_rollout_kernel' 2 2
cap nv_step     nv_step     'nv_step_kernel'      7 1
cap net         net         'net_jit_rollout'     1 1
cap net_step    net_step    'net_jit_step'        11 1
cap net64       net64       'net_jit_step|net_obs_kernel' 12 2
ls -la gpurun_out/*.ncu-rep
