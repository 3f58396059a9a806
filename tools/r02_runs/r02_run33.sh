#!/bin/bash
mkdir -p gpurun_out
for r in 1 0; do
  echo "== RING32=$r"
  ORGYM_NET_RING32=$r INFO=0 python tools/net64_quick.py 2>&1 | grep "net64 step"
  ORGYM_NET_RING32=$r INFO=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"net_obs|net_jit_step" -s 8 -c 2 python tools/net64_quick.py 2>&1 | grep -E "net_obs_kernel|net_jit_step|duration|dram"
done
python -m pytest tests/test_netinv_gpu.py tests/test_random_configs_gpu.py tests/test_canary_gpu.py -m gpu -x -q 2>&1 | tail -3
