#!/bin/bash
# final bench lines (instruction counts now match the sources) + full GPU suite + smoke
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -5 gpurun_out/r02_bench_1gpu.err; }
for wl in invmgmt_backlog invmgmt_random newsvendor netinv netinv64_mlp; do
  python bench.py --workload $wl > gpurun_out/r02_bench_1gpu_$wl.json 2> gpurun_out/r02_bench_1gpu_$wl.err || tail -3 gpurun_out/r02_bench_1gpu_$wl.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gputests_final.log
tail -3 gpurun_out/r02_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print('value %.4g e2e %.4g frac %s'%(d['value'], d['e2e']['value'], d['roofline']['frac']))
for k,v in d['other_configs'].items(): print(k, '%.4g'%v['env_steps_per_s'], (v.get('roofline') or {}).get('frac'))
"
