#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_newsvendor_gpu.py tests/test_random_configs_gpu.py tests/test_netinv_gpu.py -m gpu -x -q > gpurun_out/r02_tests14.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_tests14.log
tail -3 gpurun_out/r02_tests14.log
python tools/bench_quick.py nv 2>&1 | grep rollout
INFO=0 python tools/net64_quick.py 2>&1 | grep step
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench 2gpu exit $?"
tail -2 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_2gpu.json"))
print("N=2 value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["with_per_episode_returns"]["value"], "n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], d["episode_stats"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload netinv64_mlp > gpurun_out/r02_bench_2gpu_mlp.json 2> gpurun_out/r02_bench_2gpu_mlp.err; echo "bench mlp 2gpu exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu_mlp.json')); print('mlp N=2', d['value'], d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference | head -c 400
