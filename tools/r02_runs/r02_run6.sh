#!/bin/bash
# round 2, GPU call 6: AOT streaming kernel tests + timing; report kernel timing; bench smoke for cfg5 graph
mkdir -p gpurun_out
python -m pytest tests/test_netinv_gpu.py tests/test_canary_gpu.py -m gpu -x -q > gpurun_out/r02_tests6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests6.log
tail -8 gpurun_out/r02_tests6.log
L=gpurun_out/r02_net64_aot.log
for info in 0 1; do echo "== INFO=$info AOT streaming kernel" >> $L; INFO=$info python tools/net64_quick.py 2>&1 | grep -E "create|step" >> $L; done
echo "== INFO=0 JIT streaming kernel (AOT=0)" >> $L; ORGYM_NET_STREAM_AOT=0 ORGYM_NET_JIT_PREFETCH=0 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "create|step" >> $L
cat $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_net64_launches_aot.csv python tools/prof_net64.py > gpurun_out/ncu.log 2>&1
grep -E "net_stream|net_obs" gpurun_out/r02_net64_launches_aot.csv | tail -4 | awk -F'","' '{print $5, $NF}'
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
print("report kernels ms", e0.elapsed_time(e1) / 10, pkg.report_to_dict(rep), "candidates", float(rep[11]))
PY
timeout 300 python bench.py --workload netinv64_mlp --steps 4 --warmup 3 > gpurun_out/r02_bench_mlp.json 2> gpurun_out/r02_bench_mlp.err; echo "bench mlp exit $?"
tail -3 gpurun_out/r02_bench_mlp.err; head -c 2500 gpurun_out/r02_bench_mlp.json
