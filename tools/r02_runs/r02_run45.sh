#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench 8gpu exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_8gpu.json')); print('N=8 value %.4g e2e %.4g (%.4g)'%(d['value'], d['e2e']['value'], d['e2e']['with_per_episode_returns']['value']), d['ms_per_step'], d['clocks'])"
