#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_invmgmt_gpu.py -m gpu -x -q -k "evaluate or report" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -5 gpurun_out/r02_bench_1gpu.err; }
for wl in invmgmt_backlog invmgmt_random newsvendor netinv netinv64_mlp; do
  python bench.py --workload $wl > gpurun_out/r02_bench_1gpu_$wl.json 2> gpurun_out/r02_bench_1gpu_$wl.err || tail -3 gpurun_out/r02_bench_1gpu_$wl.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print('value %.4g e2e %.4g (%.4g / %.4g) frac %s'%(d['value'], d['e2e']['value'], d['e2e']['with_per_episode_returns']['value'], d['e2e']['with_per_episode_statistics']['value'], d['roofline']['frac']))
for k,v in d['other_configs'].items(): print(k, '%.4g'%v['env_steps_per_s'], (v.get('roofline') or {}).get('frac'))
for w in ('invmgmt_backlog','invmgmt_random','newsvendor','netinv','netinv64_mlp'):
    x=json.load(open('gpurun_out/r02_bench_1gpu_%s.json'%w)); print(w, '%.4g e2e %.4g'%(x['value'], x['e2e']['value']), (x.get('roofline') or {}).get('frac'))
"
