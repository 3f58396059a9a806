#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench 2gpu exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print('N=2 value %.4g e2e %.4g (%.4g)'%(d['value'], d['e2e']['value'], d['e2e']['with_per_episode_returns']['value']), d['ms_per_step'])"
