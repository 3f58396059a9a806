#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_invmgmt_gpu.py -m gpu -x -q -k "report or evaluate" 2>&1 | tail -2
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import or_gym_inventory_b200 as pkg
env = pkg.InvManagementLostSalesEnv(num_envs=1 << 24, device="cuda:0")
out = env.rollout("base_stock", seed=5000, safety_factor=1.0, want=("ep_return", "stats32"))
rep, scr = pkg.evaluation_report_device(out, 30)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): rep, scr = pkg.evaluation_report_device(out, 30, report=rep, scratch=scr)
e1.record(); torch.cuda.synchronize()
print("report kernels ms", e0.elapsed_time(e1) / 20, "candidates", float(rep[11]), pkg.report_to_dict(rep)["MedianReward"])
PY
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print('value %.4g e2e %.4g frac %s'%(d['value'], d['e2e']['value'], d['roofline']['frac']))"
