#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --workload netinv64_mlp --no-extras > gpurun_out/r02_bench_1gpu_netinv64_mlp.json 2> gpurun_out/r02_bench_mlp.err; echo "mlp exit $?"
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_netinv64_mlp.json'))
print('mlp value %.4g e2e %.4g'%(d['value'], d['e2e']['value']), json.dumps(d['roofline'])[:700])"
python tools/bench_quick.py net 2>&1 | tail -4
