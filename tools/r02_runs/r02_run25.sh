#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_netinv_gpu.py tests/test_random_configs_gpu.py tests/test_canary_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_quick.py net 2>&1 | tail -4
ORGYM_NET_JIT_ONEPASS=0 python tools/bench_quick.py net 2>&1 | tail -2
