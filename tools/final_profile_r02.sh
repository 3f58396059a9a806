#!/bin/bash
# Round 2, one gpurun call: bench lines for every workload + the reference arm, the ncu launch list of the bench
# command, and one `ncu --set full` capture per kernel the rooflines quote.  Every ncu command runs only after the same
# command has exited 0 without ncu.  Outputs land in gpurun_out/ (summaries are made on the CPU box: tools/ncu_summary.py).
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -5 gpurun_out/r02_bench_1gpu.err; exit 1; }
for wl in invmgmt_backlog invmgmt_random newsvendor netinv netinv64_mlp; do
  python bench.py --workload $wl > gpurun_out/r02_bench_1gpu_$wl.json 2> gpurun_out/r02_bench_1gpu_$wl.err || tail -3 gpurun_out/r02_bench_1gpu_$wl.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null
python bench.py --steps 2 --warmup 3 --inner 2 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
      python bench.py --steps 2 --warmup 3 --inner 2 > gpurun_out/ncu_bench.log 2>&1
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
cap nv          nv          'nv_level_kernel|nv_rollout_kernel' 2 2
cap nv_step     nv_step     'nv_step_kernel'      7 1
cap net         net         'net_jit_rollout'     1 1
cap net_step    net_step    'net_jit_step'        11 1
cap net64       net64       'net_jit_step|net_obs_kernel' 12 2
ls -la gpurun_out | tail -30
