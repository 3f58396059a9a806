#!/bin/bash
# One gpurun call: benches (all workloads), launch list, and one `ncu --set full` capture per family.
# Every ncu command runs only after the same command has exited 0 without ncu.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_invmgmt.json 2> gpurun_out/bench_invmgmt.err || exit 1
python bench.py --workload newsvendor --steps 20 --warmup 3 > gpurun_out/bench_newsvendor.json 2> gpurun_out/bench_newsvendor.err
python bench.py --workload netinv --steps 20 --warmup 3 > gpurun_out/bench_netinv.json 2> gpurun_out/bench_netinv.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 5 --warmup 3 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv \
      python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
for fam in inv nv net; do
  python tools/prof_quick.py $fam > /dev/null 2>&1 || continue
  case $fam in
    inv) rx='inv_step_kernel|inv_jit_rollout|inv_rollout_kernel'; skip=13; cnt=2;;   # last step launch + first rollout
    nv)  rx='nv_step_kernel|nv_rollout_kernel';  skip=7;  cnt=2;;
    net) rx='net_jit_step|net_jit_rollout';       skip=11; cnt=2;;
  esac
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o gpurun_out/final_$fam \
      python tools/prof_quick.py $fam > gpurun_out/ncu_$fam.log 2>&1
done
python tools/prof_net64.py > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:net_jit_step -s 6 -c 1 -o gpurun_out/final_net64 \
      python tools/prof_net64.py > gpurun_out/ncu_net64.log 2>&1
ls -la gpurun_out
