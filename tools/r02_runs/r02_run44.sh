#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -5 gpurun_out/r02_bench_1gpu.err; exit 1; }
for wl in invmgmt_backlog invmgmt_random; do
  python bench.py --workload $wl > gpurun_out/r02_bench_1gpu_$wl.json 2> gpurun_out/r02_bench_1gpu_$wl.err || tail -3 gpurun_out/r02_bench_1gpu_$wl.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null
python bench.py --steps 2 --warmup 3 --inner 2 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
      python bench.py --steps 2 --warmup 3 --inner 2 > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import json
for n in ("", "_invmgmt_backlog", "_invmgmt_random"):
    d = json.load(open(f"gpurun_out/r02_bench_1gpu{n}.json"))
    print(n or "invmgmt", "value %.4g e2e %.4g" % (d["value"], (d.get("e2e") or {}).get("value", 0)), "frac %.3f" % d["roofline"]["frac"], d["roofline"].get("kernel_ms"))
d = json.load(open("gpurun_out/r02_bench_1gpu.json"))
for k, v in d["other_configs"].items():
    print(" other", k, "%.4g" % v["env_steps_per_s"], "frac %.3f" % v["roofline"]["frac"])
print(json.load(open("gpurun_out/r02_bench_reference_arm.json"))["value"])
PY
