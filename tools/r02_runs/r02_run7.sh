#!/bin/bash
# round 2, GPU call 7: AOT streaming kernel with L2 prefetch + batched loads; report kernel timing; tests
mkdir -p gpurun_out
python -m pytest tests/test_netinv_gpu.py tests/test_canary_gpu.py tests/test_invmgmt_gpu.py -m gpu -x -q -k "not specialised_rollout_matches" > gpurun_out/r02_tests7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests7.log
tail -8 gpurun_out/r02_tests7.log
L=gpurun_out/r02_net64_aot2.log
for info in 0 1; do echo "== INFO=$info AOT streaming kernel v2" >> $L; INFO=$info python tools/net64_quick.py 2>&1 | grep -E "step" >> $L; done
cat $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_net64_launches_aot2.csv python tools/prof_net64.py > gpurun_out/ncu.log 2>&1
grep -E "net_stream|net_obs" gpurun_out/r02_net64_launches_aot2.csv | tail -4 | awk -F'","' '{print $5, $NF}'
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import or_gym_inventory_b200 as pkg
env = pkg.InvManagementLostSalesEnv(num_envs=1 << 24, device="cuda:0")
out = env.rollout("base_stock", seed=5000, safety_factor=1.0, want=("ep_return", "stats32"))
rep, scr = pkg.evaluation_report_device(out, 30)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): rep, scr = pkg.evaluation_report_device(out, 30, report=rep, scratch=scr)
e1.record(); torch.cuda.synchronize()
print("report kernels ms", e0.elapsed_time(e1) / 10, "candidates", float(rep[11]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_report_launches.csv python - > /dev/null 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
import or_gym_inventory_b200 as pkg
env = pkg.InvManagementLostSalesEnv(num_envs=1 << 24, device="cuda:0")
out = env.rollout("base_stock", seed=5000, safety_factor=1.0, want=("ep_return", "stats32"))
rep, scr = pkg.evaluation_report_device(out, 30)
rep, scr = pkg.evaluation_report_device(out, 30, report=rep, scratch=scr)
torch.cuda.synchronize()
PY
grep -E "report_" gpurun_out/r02_report_launches.csv | tail -17 | awk -F'","' '{print substr($5,1,40), $NF}'
python tools/bench_quick.py nv 2>&1 | grep rollout
