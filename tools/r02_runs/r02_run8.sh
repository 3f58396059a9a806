#!/bin/bash
# round 2, GPU call 8: obs kernel with L2 prefetch; report timing; newsvendor with smem reciprocal table; full GPU suite; bench
mkdir -p gpurun_out
L=gpurun_out/r02_net64_v4.log
for info in 0 1; do echo "== INFO=$info JIT stream + obs v3 + L2 prefetch" >> $L; ORGYM_NET_JIT_PREFETCH=0 INFO=$info python tools/net64_quick.py 2>&1 | grep -E "step" >> $L; done
echo "== INFO=0 THREADS=256" >> $L; ORGYM_NET_JIT_PREFETCH=0 ORGYM_NET_JIT_THREADS=256 INFO=0 python tools/net64_quick.py 2>&1 | grep -E "step" >> $L
cat $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_net64_launches_v4.csv python tools/prof_net64.py > gpurun_out/ncu.log 2>&1
grep -E "net_jit_step|net_obs" gpurun_out/r02_net64_launches_v4.csv | tail -4 | awk -F'","' '{print $5, $NF}'
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
python tools/bench_quick.py nv 2>&1 | grep rollout
python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests8.log
tail -6 gpurun_out/r02_tests8.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo "bench exit $?"
tail -3 gpurun_out/r02_bench_full.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_full.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "steps/ms", d["ms_per_step"], "clocks", d.get("clocks"))
print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "kernel_ms", "frac", "note")})
print("step_api", d["roofline_step_api"]["frac"], d["e2e_step_api"]["value"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("python_reference", {}).get("all_cores_pool"))
for k, v in d.get("other_configs", {}).items():
    print(k, v.get("env_steps_per_s"), v.get("ms_per_rollout"), (v.get("roofline") or {}).get("frac"), v.get("error"))
PY
